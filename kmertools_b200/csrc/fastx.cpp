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

ByteSource::~ByteSource() {
    if (gz_) gzclose((gzFile)gz_);
    else if (own_fd_ && fd_ >= 0) ::close(fd_);
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
        gz_ = gzdopen(fd_, "rb");
        if (!gz_) {
            if (err) *err = "Unable to open: " + path;
            return false;
        }
        gzbuffer((gzFile)gz_, 1 << 20);
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
        const int r = gzread((gzFile)gz_, p + got, (unsigned)std::min<size_t>(n - got, 1u << 30));
        if (r < 0) return -1;
        return (long)(got + r);
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
    while ((size_t)added < max_records) {
        rec_start = w = *used;
        overflow = false;
        if (fmt_ == SeqFormat::Fasta) {
            if (!have_header_) {
                if (!next_line(&lp, &ln)) break;  // end of input
                if (ln == 0 || lp[0] != '>') {
                    if (trim_end(lp, ln) == 0) continue;  // tolerate blank lines between records
                    err_ = "Expected > at record start.";
                    return -1;
                }
                have_header_ = true;
            }
            bool more = false;
            while (next_line(&lp, &ln)) {
                if (ln > 0 && lp[0] == '>') { more = true; break; }
                put(lp, trim_end(lp, ln));
            }
            have_header_ = more;  // the '>' line just consumed opens the next record
        } else {
            if (!next_line(&lp, &ln)) break;
            if (ln == 0 || lp[0] != '@') {
                if (trim_end(lp, ln) == 0) continue;
                err_ = "Expected @ at record start.";
                return -1;
            }
            size_t lines = 0;
            bool plus = false;
            while (next_line(&lp, &ln)) {
                if (ln > 0 && lp[0] == '+') { plus = true; break; }
                put(lp, trim_end(lp, ln));
                ++lines;
            }
            if (!plus) {
                err_ = "Incomplete record. Each FastQ record has to consist of 4 lines: header, sequence, separator and qualities.";
                return -1;
            }
            for (size_t i = 0; i < lines; ++i)
                if (!next_line(&lp, &ln)) break;  // quality lines are skipped
        }
        if (!commit()) return added;
    }
    return added;
}

}  // namespace ktb
