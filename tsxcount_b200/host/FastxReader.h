// FastxReader.h — host feeder: FASTQ/FASTA(+gz) records -> batches of concatenated sequences.
//
// Replaces FASTXreader<FASTQEntry> (src/fastxutils/FastXReader.h:118-478 of mjoppich/tsxCount) on the
// feeder side of the boundary, with the same record semantics:
//   - lines of length 0 are skipped wherever they occur (:367-368);
//   - a record is getLinesRequired() consecutive non-empty lines: 4 for FASTQ (:95), 2 for FASTA (:112);
//     the sequence is the 2nd line (:70, :108); a trailing incomplete record is dropped (:242);
//   - a file whose name ends in ".gz" is inflated with zlib (:185-190, :387-440).
// Instead of a vector of entry objects with three std::string copies per record (:239-253) the reader
// appends the sequences of a batch to one buffer plus an offsets array — the layout tsxc_pack_reads takes.
#pragma once

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

class FastxReader {
public:
    explicit FastxReader(const std::string& path, int lines_per_record = 4) : m_lines(lines_per_record) {
        // gzopen reads plain files transparently, so one code path serves both (the reference sniffs ".gz")
        m_file = gzopen(path.c_str(), "rb");
        if (!m_file) throw std::runtime_error("FastxReader: cannot open " + path);
        gzbuffer(m_file, 1 << 20);
        m_buf.resize(1 << 20);
    }
    ~FastxReader() { if (m_file) gzclose(m_file); }
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;

    bool hasNext() { return fill() ; }

    // Appends up to max_reads sequences to `bases`; offsets gets n+1 entries (offsets[0] == 0).
    // Returns the number of reads delivered (0 at end of file).
    size_t nextBatch(size_t max_reads, std::string& bases, std::vector<uint64_t>& offsets) {
        bases.clear();
        offsets.assign(1, 0);
        std::string line;
        while (offsets.size() - 1 < max_reads && getLine(line)) {
            if (line.empty()) continue;
            if (m_in_record == 1) m_seq.swap(line);
            if (++m_in_record == m_lines) {
                bases.append(m_seq);
                offsets.push_back(bases.size());
                m_in_record = 0;
            }
        }
        return offsets.size() - 1;
    }

private:
    bool fill() {
        if (m_pos < m_len) return true;
        if (m_eof) return false;
        const int n = gzread(m_file, m_buf.data(), (unsigned)m_buf.size());
        if (n <= 0) { m_eof = true; return false; }
        m_pos = 0; m_len = (size_t)n;
        return true;
    }
    // std::getline semantics: strips '\n' only
    bool getLine(std::string& out) {
        out.clear();
        bool any = false;
        while (fill()) {
            any = true;
            const char* b = m_buf.data() + m_pos;
            const char* e = m_buf.data() + m_len;
            const char* nl = b;
            while (nl < e && *nl != '\n') ++nl;
            out.append(b, nl);
            m_pos = (size_t)(nl - m_buf.data());
            if (nl < e) { ++m_pos; return true; }
        }
        return any;
    }

    gzFile m_file = nullptr;
    int m_lines;
    int m_in_record = 0;
    std::string m_seq;
    std::vector<char> m_buf;
    size_t m_pos = 0, m_len = 0;
    bool m_eof = false;
};
