// FastxReader.h — host feeder: FASTQ/FASTA(+gz) records -> batches of concatenated sequences.
//
// Replaces FASTXreader<FASTQEntry> (src/fastxutils/FastXReader.h:118-478 of mjoppich/tsxCount) on the
// feeder side of the boundary, with the same record semantics:
//   - lines of length 0 are skipped wherever they occur (:367-368);
//   - a record is getLinesRequired() consecutive non-empty lines: 4 for FASTQ (:95), 2 for FASTA (:112);
//     the sequence is the 2nd line (:70, :108); a trailing incomplete record is dropped (:242);
//   - gzip input is inflated with zlib (:387-440; the reference sniffs the ".gz" suffix, :185-190 — here the
//     gzip magic bytes decide, so a mis-named file still works).  Files written by bgzip (BGZF: independent gzip
//     members of at most 64 KiB, each announcing its size in a 'BC' extra field) are inflated by several threads,
//     block by block, with the CRC of every block checked (BgzfSource below); any other gzip stream is a single
//     deflate stream and goes through gzread on one thread like in the reference.
// Beyond the reference (opt-in, lines_per_record = 0 = auto-detect): a file whose first non-empty byte is '>' is read
// as multi-line FASTA — a record is a '>' header plus all following lines up to the next header, lower-case
// (soft-masked) bases are folded to upper case, and a sequence longer than piece_len bases is delivered as pieces
// that overlap by `overlap` (= k-1) bases, which preserves the multiset of k-mers exactly.
// Instead of a vector of entry objects with three std::string copies per record (:239-253) the reader scans
// lines in place in a large block buffer (memchr) and appends only the sequence lines to one buffer plus an
// offsets array — the layout tsxc_pack_reads takes.  ~0.8 GB/s of FASTQ text per thread.
#pragma once

#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

// Parallel inflate of a BGZF file (the blocked gzip variant bgzip / htslib write).  fill() reads a chunk of the file,
// walks the block headers (each carries its own compressed size), and inflates the blocks of the chunk with n threads
// into one contiguous buffer (every block also carries its uncompressed size, so the output offsets are known up
// front).  Damage (bad header, inflate error, size or CRC mismatch, a file that ends inside a block) throws.
class BgzfSource {
public:
    // true iff the file starts with a gzip member that has the BGZF 'BC' extra subfield
    static bool isBgzf(int fd) {
        unsigned char h[18];
        if (::pread(fd, h, sizeof h, 0) != (ssize_t)sizeof h) return false;
        return blockSize(h, sizeof h) > 0;
    }
    BgzfSource(int fd, int threads, size_t chunk_bytes = 32u << 20)
        : m_fd(fd), m_threads(std::max(1, threads)), m_in(std::max<size_t>(chunk_bytes, 1u << 17)) {}

    // Copies up to n inflated bytes to dst; 0 = end of file.
    size_t read(char* dst, size_t n) {
        while (m_out_pos == m_out.size()) {
            if (m_done) return 0;
            fill();
        }
        const size_t take = std::min(n, m_out.size() - m_out_pos);
        std::memcpy(dst, m_out.data() + m_out_pos, take);
        m_out_pos += take;
        return take;
    }

private:
    struct Block { size_t in_off, in_len, out_off, out_len; };
    // total size of the BGZF block that starts at h (0: not a BGZF block header / header incomplete)
    static size_t blockSize(const unsigned char* h, size_t avail) {
        if (avail < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
        const size_t xlen = h[10] | ((size_t)h[11] << 8);
        if (avail < 12 + xlen) return 0;
        for (size_t p = 12; p + 4 <= 12 + xlen;) {
            const size_t slen = h[p + 2] | ((size_t)h[p + 3] << 8);
            if (h[p] == 'B' && h[p + 1] == 'C' && slen == 2 && p + 6 <= 12 + xlen) return (size_t)(h[p + 4] | (h[p + 5] << 8)) + 1;
            p += 4 + slen;
        }
        return 0;
    }
    void fill() {
        m_out.clear(); m_out_pos = 0;
        size_t got = 0;
        while (got < m_in.size()) {
            const ssize_t r = ::pread(m_fd, m_in.data() + got, m_in.size() - got, (off_t)(m_off + got));
            if (r < 0) { if (errno == EINTR) continue; throw std::runtime_error(std::string("FastxReader: read failed: ") + std::strerror(errno)); }
            if (r == 0) break;
            got += (size_t)r;
        }
        if (got == 0) { m_done = true; return; }
        std::vector<Block> blocks;
        size_t pos = 0, out = 0;
        while (pos < got) {
            const unsigned char* h = (const unsigned char*)m_in.data() + pos;
            const size_t bs = blockSize(h, got - pos);
            if (bs == 0 || bs < 26) {
                if (got - pos < 18 + 65536 && got == m_in.size()) break;      // header cut by the end of the chunk
                throw std::runtime_error("FastxReader: gzip stream damaged or truncated: not a BGZF block");
            }
            if (pos + bs > got) {
                if (got == m_in.size()) break;                                // block continues in the next chunk
                throw std::runtime_error("FastxReader: gzip stream damaged or truncated: file ends inside a BGZF block");
            }
            const size_t xlen = h[10] | ((size_t)h[11] << 8);
            const unsigned char* tail = h + bs - 8;
            const size_t isize = tail[4] | ((size_t)tail[5] << 8) | ((size_t)tail[6] << 16) | ((size_t)tail[7] << 24);
            if (isize > 65536) throw std::runtime_error("FastxReader: gzip stream damaged or truncated: BGZF block larger than 64 KiB");
            blocks.push_back({pos + 12 + xlen, bs - 12 - xlen - 8, out, isize});
            out += isize;
            pos += bs;
        }
        if (pos == 0) throw std::runtime_error("FastxReader: gzip stream damaged or truncated: BGZF block exceeds the read buffer");
        m_off += pos;
        m_out.resize(out);
        const int nt = (int)std::min<size_t>((size_t)m_threads, std::max<size_t>(1, blocks.size() / 4));
        std::vector<std::string> errors((size_t)nt);
        auto work = [&](int t) {
            z_stream zs;
            std::memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { errors[(size_t)t] = "inflateInit2 failed"; return; }
            for (size_t i = (size_t)t; i < blocks.size() && errors[(size_t)t].empty(); i += (size_t)nt) {
                const Block& b = blocks[i];
                zs.next_in = (Bytef*)(m_in.data() + b.in_off); zs.avail_in = (uInt)b.in_len;
                zs.next_out = (Bytef*)(m_out.data() + b.out_off); zs.avail_out = (uInt)b.out_len;
                const int rc = b.out_len || b.in_len > 2 ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
                const unsigned char* tail = (const unsigned char*)m_in.data() + b.in_off + b.in_len;
                const uint32_t want_crc = tail[0] | ((uint32_t)tail[1] << 8) | ((uint32_t)tail[2] << 16) | ((uint32_t)tail[3] << 24);
                if (rc != Z_STREAM_END || zs.avail_out != 0) errors[(size_t)t] = "inflate error in a BGZF block";
                else if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)(m_out.data() + b.out_off), (uInt)b.out_len) != want_crc)
                    errors[(size_t)t] = "CRC mismatch in a BGZF block";
                inflateReset(&zs);
            }
            inflateEnd(&zs);
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        for (const auto& e : errors) if (!e.empty()) throw std::runtime_error("FastxReader: gzip stream damaged or truncated: " + e);
    }

    int m_fd;
    int m_threads;
    std::vector<char> m_in;
    std::vector<char> m_out;
    size_t m_out_pos = 0;
    uint64_t m_off = 0;
    bool m_done = false;
};

class FastxReader {
public:
    // Whole file, or (plain files only) the records that START in the byte range [range_begin, range_end):
    // several readers on disjoint ranges of one file deliver every record exactly once.  A reader that does not
    // start at byte 0 has to find a record boundary by itself: for FASTQ the first line that begins with '@'
    // and whose second next non-empty line begins with '+' (a quality line may begin with '@', but then the
    // line two further down is a sequence); for FASTA the first line beginning with '>'.  This is exact for
    // well-formed files only, so ranges are opt-in (the CLI's --readers); the default reads sequentially with
    // the reference's tolerant semantics.
    // inflate_threads > 1: a BGZF (bgzip) input is inflated by that many threads; ignored for everything else.
    explicit FastxReader(const std::string& path, int lines_per_record = 4, size_t block_bytes = 8u << 20,
                         uint64_t range_begin = 0, uint64_t range_end = ~0ULL, int inflate_threads = 1)
        : m_lines(lines_per_record), m_buf(block_bytes), m_range_end(range_end) {
        m_fd = ::open(path.c_str(), O_RDONLY);
        if (m_fd < 0) throw std::runtime_error("FastxReader: cannot open " + path);
        unsigned char magic[2] = {0, 0};
        const ssize_t got = ::pread(m_fd, magic, 2, 0);
        auto die = [&](const std::string& msg) {       // a throwing constructor runs no destructor
            delete m_bgzf; m_bgzf = nullptr;
            if (m_gz) gzclose(m_gz); else ::close(m_fd);
            throw std::runtime_error("FastxReader: " + msg);
        };
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            if (range_begin != 0 || range_end != ~0ULL) die("byte ranges need a plain file");
            if (inflate_threads > 1 && BgzfSource::isBgzf(m_fd)) {
                m_bgzf = new BgzfSource(m_fd, inflate_threads);
            } else {
                m_gz = gzdopen(m_fd, "rb");
                if (!m_gz) die("gzdopen failed for " + path);
                gzbuffer(m_gz, 1 << 20);
            }
        }
        if (m_lines == 0) {                            // auto: '>' -> multi-line FASTA, anything else -> FASTQ
            m_multi_fasta = (range_begin == 0 ? firstByte() : sniff(path)) == '>';
            m_lines = m_multi_fasta ? 2 : 4;
            if (m_multi_fasta && (range_begin != 0 || range_end != ~0ULL)) die("byte ranges need FASTQ input");
        }
        if (!m_gz && !m_bgzf && range_begin > 0) {
            // start one byte early: if that byte is '\n', range_begin is the start of a line
            if (::lseek(m_fd, (off_t)(range_begin - 1), SEEK_SET) < 0) die("seek failed");
            m_pos = m_end = 0; m_eof = false;          // auto-detection may have buffered the start of the file
            m_file_off = range_begin - 1;
            const char* l; size_t n; uint64_t off;
            nextLine(l, n, off);                       // the (rest of the) line we landed in belongs to the previous range
            syncToRecord();
        }
    }

    // Multi-line FASTA only: sequences longer than piece_len are split into pieces overlapping by `overlap` bases.
    void setFastaSplit(size_t piece_len, size_t overlap) {
        if (piece_len <= overlap) throw std::runtime_error("FastxReader: piece length must exceed the overlap");
        m_piece_len = piece_len; m_overlap = overlap;
    }
    bool isMultiFasta() const { return m_multi_fasta; }
    // '@', '>' or 0 (empty / unreadable): first non-empty byte of the (possibly gzip) file
    static char sniff(const std::string& path) {
        try { FastxReader r(path, 4, 1u << 16); return r.firstByte(); } catch (...) { return 0; }
    }

    static bool isGzip(const std::string& path) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        unsigned char magic[2] = {0, 0};
        const ssize_t got = ::pread(fd, magic, 2, 0);
        ::close(fd);
        return got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    }
    bool parallelInflate() const { return m_bgzf != nullptr; }
    ~FastxReader() {
        delete m_bgzf;
        if (m_gz) gzclose(m_gz);        // closes the descriptor too
        else if (m_fd >= 0) ::close(m_fd);
    }
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;

    // Appends up to max_reads sequences to `bases`; offsets gets n+1 entries (offsets[0] == 0).
    // Returns the number of reads delivered (0 at end of file).
    size_t nextBatch(size_t max_reads, std::string& bases, std::vector<uint64_t>& offsets) {
        if (m_multi_fasta) return nextBatchFasta(max_reads, bases, offsets);
        bases.clear();
        offsets.assign(1, 0);
        while (offsets.size() - 1 < max_reads && !m_range_done) {
            const char* line;
            size_t len;
            uint64_t off;
            if (!nextLine(line, len, off)) break;
            if (len == 0) continue;
            if (m_in_record == 0 && off >= m_range_end) { m_range_done = true; break; }   // starts in the next range
            if (m_in_record == 1) bases.append(line, len);   // the record may still turn out incomplete at EOF
            if (++m_in_record == m_lines) {
                offsets.push_back(bases.size());
                m_in_record = 0;
            }
        }
        bases.resize(offsets.back());                          // drop the sequence of a trailing incomplete record
        return offsets.size() - 1;
    }

private:
    // Multi-line FASTA.  The piece under construction is the tail of `bases` after offsets.back(); when a batch
    // ends with an open piece, the piece moves to m_carry and continues in the next batch.
    size_t nextBatchFasta(size_t max_reads, std::string& bases, std::vector<uint64_t>& offsets) {
        const size_t kMaxBatchBases = 64u << 20;
        bases.assign(m_carry);
        m_carry.clear();
        offsets.assign(1, 0);
        auto close_piece = [&] {                       // end of a record: deliver what is left if it has new bases
            if (m_fresh) offsets.push_back(bases.size()); else bases.resize(offsets.back());
            m_fresh = 0;
        };
        while (offsets.size() - 1 < max_reads && offsets.back() < kMaxBatchBases) {
            const char* line;
            size_t len;
            uint64_t off;
            if (!nextLine(line, len, off)) { close_piece(); return offsets.size() - 1; }
            if (len == 0) continue;
            if (line[0] == '>') { close_piece(); continue; }
            for (size_t consumed = 0; consumed < len;) {
                const size_t cur = bases.size() - offsets.back();
                const size_t take = std::min(len - consumed, m_piece_len - cur);
                const size_t at = bases.size();
                bases.append(line + consumed, take);
                char* dst = &bases[at];
                for (size_t i = 0; i < take; ++i) { const char c = dst[i]; dst[i] = (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
                consumed += take;
                m_fresh += take;
                if (bases.size() - offsets.back() == m_piece_len) {     // piece full: the next one starts with its last k-1 bases
                    const size_t end = bases.size();
                    offsets.push_back(end);
                    const std::string lap = bases.substr(end - m_overlap, m_overlap);
                    bases.append(lap);
                    m_fresh = 0;
                }
            }
        }
        // batch full: an open piece continues in the next batch
        m_carry.assign(bases, offsets.back(), std::string::npos);
        bases.resize(offsets.back());
        return offsets.size() - 1;
    }
    char firstByte() {
        for (;;) {
            for (size_t i = m_pos; i < m_end; ++i) if (m_buf[i] != '\n') return m_buf[i];
            if (m_eof) return 0;
            m_file_off += m_end; m_pos = m_end = 0;    // only newlines so far: drop them
            const size_t got = readSome(m_buf.data(), m_buf.size());
            if (got == 0) m_eof = true;
            m_end = got;
        }
    }
    // 0 = clean end of file.  I/O errors and damaged / truncated gzip streams throw: a partial count with exit
    // code 0 would look like a result.
    size_t readSome(char* dst, size_t n) {
        if (m_bgzf) return m_bgzf->read(dst, n);
        if (m_gz) {
            const int r = gzread(m_gz, dst, (unsigned)n);
            if (r > 0) return (size_t)r;
            int errnum = Z_OK;
            const char* msg = gzerror(m_gz, &errnum);
            if (r < 0 || (errnum != Z_OK && errnum != Z_STREAM_END))
                throw std::runtime_error(std::string("FastxReader: gzip stream damaged or truncated: ") + (msg ? msg : "?"));
            return 0;
        }
        for (;;) {
            const ssize_t r = ::read(m_fd, dst, n);
            if (r >= 0) return (size_t)r;
            if (errno == EINTR) continue;
            throw std::runtime_error(std::string("FastxReader: read failed: ") + std::strerror(errno));
        }
    }
    // Positions the reader on the first record boundary at or after the current line (see the constructor).
    void syncToRecord() {
        struct L { uint64_t off; char first; };
        std::vector<L> win;                            // non-empty lines seen so far, with their file offsets
        const char first_char = m_lines == 4 ? '@' : '>';
        for (;;) {
            const char* l; size_t n; uint64_t off;
            if (!nextLine(l, n, off)) { m_range_done = true; return; }
            if (n == 0) continue;
            win.push_back({off, l[0]});
            const size_t need = m_lines == 4 ? 3 : 1;  // FASTQ: header, sequence, '+'
            while (win.size() >= need) {
                const bool ok = win[0].first == first_char && (m_lines != 4 || win[2].first == '+');
                if (ok) { rewindTo(win[0].off); return; }
                win.erase(win.begin());
            }
        }
    }
    // Re-reads from an absolute file offset (used once, after syncToRecord looked ahead).
    void rewindTo(uint64_t off) {
        if (::lseek(m_fd, (off_t)off, SEEK_SET) < 0) throw std::runtime_error("FastxReader: seek failed");
        m_file_off = off; m_pos = m_end = 0; m_eof = false;
    }

    // Next line as a view into the block buffer (valid until the next call); std::getline semantics: the
    // terminating '\n' is stripped, nothing else; a last line without '\n' is delivered.  off = file offset of
    // the line's first byte (plain files).
    bool nextLine(const char*& line, size_t& len, uint64_t& off) {
        for (;;) {
            const char* b = m_buf.data() + m_pos;
            const char* nl = m_end > m_pos ? (const char*)std::memchr(b, '\n', m_end - m_pos) : nullptr;
            if (nl) {
                line = b; len = (size_t)(nl - b); off = m_file_off + m_pos;
                m_pos += len + 1;
                return true;
            }
            if (m_eof) {
                if (m_end == m_pos) return false;
                line = b; len = m_end - m_pos; off = m_file_off + m_pos;
                m_pos = m_end;
                return true;
            }
            // no complete line left: move the tail to the front and refill
            const size_t tail = m_end - m_pos;
            if (tail && m_pos) std::memmove(m_buf.data(), b, tail);
            m_file_off += m_pos;
            m_pos = 0; m_end = tail;
            if (m_end == m_buf.size()) m_buf.resize(m_buf.size() * 2);   // a line longer than the block
            const size_t got = readSome(m_buf.data() + m_end, m_buf.size() - m_end);
            if (got == 0) m_eof = true;
            m_end += got;
        }
    }

    int m_fd = -1;
    gzFile m_gz = nullptr;
    BgzfSource* m_bgzf = nullptr;
    int m_lines;
    int m_in_record = 0;
    std::vector<char> m_buf;
    size_t m_pos = 0, m_end = 0;
    bool m_eof = false;
    uint64_t m_file_off = 0;       // file offset of m_buf[0]
    uint64_t m_range_end = ~0ULL;
    bool m_range_done = false;
    bool m_multi_fasta = false;
    size_t m_piece_len = 1u << 20, m_overlap = 0, m_fresh = 0;
    std::string m_carry;
};
