// fastx.cpp — see fastx.h
#include "fastx.h"

#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

namespace ktb {

static bool ends_with(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

bool format_from_path(const std::string &path_in, SeqFormat *out) {
    std::string path = path_in;
    if (ends_with(path, ".gz")) {  // trim_end_matches(".gz") removes every trailing ".gz"
        while (ends_with(path, ".gz")) path.resize(path.size() - 3);
    }
    if (ends_with(path, ".fq") || ends_with(path, ".fastq")) { *out = SeqFormat::Fastq; return true; }
    if (ends_with(path, ".fasta") || ends_with(path, ".fa") || ends_with(path, ".fna")) { *out = SeqFormat::Fasta; return true; }
    return false;
}

// gzip input: flate2::read::GzDecoder (ktio/src/seq.rs:149) decodes the FIRST member of the file and reports end of
// stream there; zlib's gzread() would run on into further members, so the stream is inflated by hand.
struct GzState {
    z_stream zs;
    std::vector<uint8_t> in;
    bool done = false;      // first member finished
    GzState() : in(1u << 20) { memset(&zs, 0, sizeof zs); }
};

ByteSource::~ByteSource() {
    if (gz_) {
        GzState *g = (GzState *)gz_;
        inflateEnd(&g->zs);
        delete g;
    }
    if (own_fd_ && fd_ >= 0) ::close(fd_);
}

bool ByteSource::open(const std::string &path, std::string *err) {
    if (path == "-") {
        fd_ = 0;
        own_fd_ = false;
        return true;
    }
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) {
        if (err) *err = "Unable to open: " + path;  // ktio/src/seq.rs:147
        return false;
    }
    own_fd_ = true;
    if (ends_with(path, ".gz")) {
        GzState *g = new GzState();
        if (inflateInit2(&g->zs, 16 + MAX_WBITS) != Z_OK) {   // 16: expect a gzip header, as GzDecoder does
            delete g;
            if (err) *err = "Unable to open: " + path;
            return false;
        }
        gz_ = g;
    }
    return true;
}

long ByteSource::read(void *buf, size_t n) {
    if (n == 0) return 0;
    uint8_t *p = (uint8_t *)buf;
    size_t got = 0;
    if (peeked_ >= 0) {
        p[0] = (uint8_t)peeked_;
        peeked_ = -2;
        got = 1;
        if (n == 1) return 1;
    }
    if (gz_) {
        GzState *g = (GzState *)gz_;
        if (g->done) return (long)got;
        g->zs.next_out = p + got;
        g->zs.avail_out = (unsigned)std::min<size_t>(n - got, 1u << 30);
        const unsigned want = g->zs.avail_out;
        while (g->zs.avail_out == want) {   // until something was produced, the member ended, or the file did
            if (g->zs.avail_in == 0) {
                ssize_t r;
                do { r = ::read(fd_, g->in.data(), g->in.size()); } while (r < 0 && errno == EINTR);
                if (r < 0) return -1;
                if (r == 0) return (want == g->zs.avail_out && !g->done) ? -1 : (long)(got + want - g->zs.avail_out);  // truncated member
                g->zs.next_in = g->in.data();
                g->zs.avail_in = (unsigned)r;
            }
            const int rc = inflate(&g->zs, Z_NO_FLUSH);
            if (rc == Z_STREAM_END) { g->done = true; break; }
            if (rc != Z_OK && rc != Z_BUF_ERROR) return -1;
        }
        return (long)(got + want - g->zs.avail_out);
    }
    for (;;) {
        const ssize_t r = ::read(fd_, p + got, n - got);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        return (long)(got + r);
    }
}

int ByteSource::peek_first_byte() {
    if (peeked_ == -2) {
        uint8_t b;
        const long r = read(&b, 1);
        peeked_ = (r == 1) ? b : -1;
    }
    return peeked_ >= 0 ? peeked_ : -1;
}

FastxParser::FastxParser(ByteSource *src, SeqFormat fmt) : src_(src), fmt_(fmt), buf_(4u << 20) {}

bool FastxParser::refill() {
    if (eof_) return false;
    pos_ = end_ = 0;
    const long r = src_->read(buf_.data(), buf_.size());
    if (r < 0) {
        err_ = "read error";
        eof_ = true;
        return false;
    }
    if (r == 0) {
        eof_ = true;
        return false;
    }
    end_ = (size_t)r;
    return true;
}

// One line without its '\n'.  The pointer stays valid until the next call.
bool FastxParser::next_line(const uint8_t **p, size_t *len) {
    line_.clear();
    bool spilled = false;
    for (;;) {
        if (pos_ == end_) {
            if (!refill()) {
                if (spilled) {  // last line without a terminator
                    *p = line_.data();
                    *len = line_.size();
                    return true;
                }
                return false;
            }
        }
        const uint8_t *s = buf_.data() + pos_;
        const uint8_t *nl = (const uint8_t *)memchr(s, '\n', end_ - pos_);
        if (nl) {
            const size_t n = (size_t)(nl - s);
            pos_ += n + 1;
            if (!spilled) {
                *p = s;
                *len = n;
            } else {
                line_.insert(line_.end(), s, s + n);
                *p = line_.data();
                *len = line_.size();
            }
            return true;
        }
        line_.insert(line_.end(), s, (const uint8_t *)(buf_.data() + end_));
        spilled = true;
        pos_ = end_;
    }
}

static inline size_t trim_end(const uint8_t *p, size_t n) {  // str::trim_end on ASCII whitespace
    while (n > 0) {
        const uint8_t c = p[n - 1];
        if (c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f') --n;
        else break;
    }
    return n;
}

long FastxParser::fill(uint8_t *bases, size_t cap, size_t *used, std::vector<uint64_t> *offsets, size_t max_records) {
    long added = 0;
    need_ = 0;
    if (!err_.empty()) return -1;
    // a complete record left over from the previous call goes first
    auto flush_pending = [&]() -> bool {
        if (!in_record_) return true;
        if (pending_.size() > cap - *used) {
            if (added == 0) need_ = pending_.size();  // caller: flush a non-empty buffer, or grow an empty one
            return false;
        }
        if (!pending_.empty()) memcpy(bases + *used, pending_.data(), pending_.size());
        *used += pending_.size();
        offsets->push_back(*used);
        pending_.clear();
        in_record_ = false;
        ++added;
        ++nrec_;
        return true;
    };
    if (!flush_pending()) return added;

    const uint8_t *lp;
    size_t ln;
    // Sequence bytes go straight into the caller's buffer; only a record that does not fit is diverted to
    // pending_ (and handed out first on the next call).
    size_t rec_start = *used, w = *used;
    bool overflow = false;
    auto put = [&](const uint8_t *src, size_t t) {
        if (!overflow && t <= cap - w) {
            memcpy(bases + w, src, t);
            w += t;
            return;
        }
        if (!overflow) {
            pending_.assign(bases + rec_start, bases + w);
            overflow = true;
        }
        pending_.insert(pending_.end(), src, src + t);
    };
    auto commit = [&]() -> bool {   // false: the record is parked in pending_, stop this batch
        if (overflow) {
            in_record_ = true;
            if (added == 0) need_ = pending_.size();
            return false;
        }
        *used = w;
        offsets->push_back(w);
        ++added;
        ++nrec_;
        return true;
    };
    while ((size_t)added < max_records && !stopped_) {
        rec_start = w = *used;
        overflow = false;
        if (fmt_ == SeqFormat::Fasta) {
            // bio::io::fasta::Reader::read (rust-bio 2.3.0): the first line of the input must start with '>' (a blank
            // line there is an error too); the sequence is every following line, trailing whitespace trimmed, up to
            // the next line that starts with '>' or the end of input.
            if (!have_header_) {
                if (!next_line(&lp, &ln)) break;  // end of input
                if (ln == 0 || lp[0] != '>') {
                    err_ = "Expected > at record start.";
                    return -1;
                }
                have_header_ = true;
                header_blank_ = trim_end(lp + 1, ln - 1) == 0;
            }
            bool more = false, next_blank = false;
            while (next_line(&lp, &ln)) {
                if (ln > 0 && lp[0] == '>') { more = true; next_blank = trim_end(lp + 1, ln - 1) == 0; break; }
                put(lp, trim_end(lp, ln));
            }
            // Records::next stops at the first EMPTY record (no id, no description, no sequence): a bare ">" line
            // directly followed by another header or the end of input ends the iteration, whatever follows.
            if (header_blank_ && !overflow && w == rec_start) {
                have_header_ = false;
                stopped_ = true;
                break;
            }
            have_header_ = more;  // the '>' line just consumed opens the next record
            header_blank_ = next_blank;
        } else {
            // bio::io::fastq::Reader::read (rust-bio 2.3.0): '@' line (anything else, a blank line included, is
            // Error::MissingAt), sequence lines up to the first line starting with '+', then exactly as many quality
            // lines as there were sequence lines; an empty quality string is Error::IncompleteRecord.
            if (!next_line(&lp, &ln)) break;
            if (ln == 0 || lp[0] != '@') {
                err_ = "Expected @ at record start.";
                return -1;
            }
            size_t lines = 0, qual = 0;
            while (next_line(&lp, &ln)) {
                if (ln > 0 && lp[0] == '+') break;
                put(lp, trim_end(lp, ln));
                ++lines;
            }
            for (size_t i = 0; i < lines; ++i) {
                if (!next_line(&lp, &ln)) break;  // quality lines are only measured
                qual += trim_end(lp, ln);
            }
            if (qual == 0) {
                err_ = "Incomplete record. Each FastQ record has to consist of 4 lines: header, sequence, separator and qualities.";
                return -1;
            }
        }
        if (!err_.empty()) return -1;   // a read / inflate error ended the input early
        if (!commit()) return added;
    }
    if (!err_.empty()) return -1;
    return added;
}

}  // namespace ktb
