// fastx.h — FASTA / FASTQ (.gz, stdin) record reader that fills offset-indexed byte buffers.
//
// Host-side feeder of the oligo path ("next" row N1 of SURVEY.md §8f).  Restates what the reference
// gets from ktio/src/seq.rs:29-155 on top of rust-bio 2.3.0 (bio::io::fasta / fastq readers, absent
// from /root/reference, pinned in Cargo.lock):
//   * format: by extension (.fa/.fasta/.fna, .fq/.fastq, optional .gz; seq.rs:29-42) or by sniffing
//     the first byte ('>' = FASTA, else FASTQ; composition/src/oligo.rs:100-104)
//   * FASTA: a record starts at a line beginning with '>'; every following line up to the next '>' is
//     sequence with trailing whitespace trimmed (multi-line records are concatenated)
//   * FASTQ: '@' header, sequence lines up to the '+' line, then as many quality lines as sequence lines
//   * ".gz" paths are inflated (zlib), "-" is stdin (seq.rs:141-155)
// The reference's own tests pin multi-line FASTA, ids, single-member gzip and a missing final newline
// (seq.rs:164-233).  Everything else follows the published source of the two readers and of flate2, restated in
// fastx.cpp next to each rule and pinned by tests/test_io_host.py: CRLF (trailing whitespace of every line is
// trimmed), blank lines (part of a FASTA sequence; Error::MissingAt where a FASTQ header is expected, also at the
// end of the file; "Expected > at record start." as the first line of a FASTA file), multi-line FASTQ (quality lines
// are counted, so '@' / '>' may start one), an empty quality string (Error::IncompleteRecord), a bare '>' record
// (ends the iteration) and multi-member gzip (flate2::read::GzDecoder stops after the first member).  The crates
// themselves cannot be executed here (no Rust toolchain), so these are restatements, not differential tests.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ktb {

enum class SeqFormat { Fasta, Fastq };

// ktio/src/seq.rs:29-42
bool format_from_path(const std::string &path, SeqFormat *out);

class ByteSource {
public:
    ByteSource() = default;
    ~ByteSource();
    ByteSource(const ByteSource &) = delete;
    ByteSource &operator=(const ByteSource &) = delete;
    bool open(const std::string &path, std::string *err);  // "-" = stdin, "*.gz" = gzip
    // up to n bytes; 0 at end of stream; -1 on error
    long read(void *buf, size_t n);
    int peek_first_byte();  // -1 when the stream is empty
private:
    int fd_ = -1;
    void *gz_ = nullptr;
    bool own_fd_ = false;
    int peeked_ = -2;  // -2 = nothing buffered
};

// Streaming parser.  fill() appends whole records to a caller-owned byte buffer (bases) and offsets
// vector until the buffer cannot take the next record or max_records is reached.
class FastxParser {
public:
    FastxParser(ByteSource *src, SeqFormat fmt);
    // Appends records to bases[0..cap) starting at *used; offsets gets one entry per record END
    // (offsets must already hold the start of the first record, i.e. *used).  Returns the number of
    // records appended, 0 at end of input, -1 on a parse error (message in error()).  A record larger
    // than the free space of an EMPTY buffer sets need_bytes() so the caller can grow it.
    long fill(uint8_t *bases, size_t cap, size_t *used, std::vector<uint64_t> *offsets, size_t max_records);
    bool eof() const { return stopped_ || (eof_ && pos_ == end_ && pending_.empty() && !in_record_ && !have_header_); }
    size_t need_bytes() const { return need_; }
    const std::string &error() const { return err_; }
    uint64_t records() const { return nrec_; }
private:
    bool next_line(const uint8_t **p, size_t *len);  // without the terminator; false at end of input
    bool refill();
    ByteSource *src_;
    SeqFormat fmt_;
    std::vector<uint8_t> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
    std::vector<uint8_t> line_;      // spill for lines crossing buffer refills
    std::vector<uint8_t> pending_;   // sequence of a record that did not fit the previous batch
    bool in_record_ = false;         // pending_ holds a COMPLETE record waiting for space
    // FASTQ state
    std::string err_;
    size_t need_ = 0;
    uint64_t nrec_ = 0;
    bool have_header_ = false;       // FASTA: a '>' line has been consumed and its record is open
    bool header_blank_ = false;      // FASTA: that line carried neither an id nor a description
    bool stopped_ = false;           // FASTA: an empty record ended the iteration (bio::io::fasta::Records::next)
};

}  // namespace ktb
