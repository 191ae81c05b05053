// Host-side I/O of the rb_align driver: FASTA/FASTQ(.gz) record reader with klib-kseq
// semantics (what the reference's loop sees, include/kseq.h:178-218), the document list
// (include/doclist.hpp) and allocation-free number formatting for the stdout grammar of
// src/rb_align.cpp:118-145.
#pragma once
#include <unistd.h>
#include <cerrno>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace rbhost {

// Record reader.  Return codes of next() follow kseq_read: >=0 sequence length, -1 end of
// file, -2 truncated quality string, -3 stream error.
class FastxReader {
  public:
    explicit FastxReader(const char* path) : fp_(gzopen(path, "r")), buf_(1 << 20) {
        if (fp_) gzbuffer(fp_, 1 << 18);
    }
    ~FastxReader() { if (fp_) gzclose(fp_); }
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;
    bool ok() const { return fp_ != nullptr; }
    // Restart between records at byte `pos` of the (uncompressed) stream.
    bool seek(size_t pos) {
        if (!fp_ || gzseek(fp_, (z_off_t) pos, SEEK_SET) < 0) return false;
        begin_ = end_ = 0;
        eof_ = err_ = false;
        last_char_ = 0;
        return true;
    }

    // Appends the record's name (up to the first whitespace) to `name` and its sequence
    // bytes to `seq` (both cleared first).
    int next(std::string& name, std::string& seq) {
        int c;
        if (last_char_ == 0) {                       // jump to the next header line
            while ((c = getc()) >= 0 && c != '>' && c != '@') {}
            if (c < 0) return c;
            last_char_ = c;
        }
        name.clear();
        seq.clear();
        int delim = 0;
        int r = until(kSpace, name, &delim, false);
        if (r < 0) return r;
        if (delim != '\n') { scratch_.clear(); until(kLine, scratch_, nullptr, false); }   // comment
        while ((c = getc()) >= 0 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;                 // skip empty lines
            seq.push_back((char) c);
            until(kLine, seq, nullptr, true);        // rest of the line
        }
        if (c == '>' || c == '@') last_char_ = c;
        if (c != '+') return (int) seq.size();       // FASTA (or end of input)
        while ((c = getc()) >= 0 && c != '\n') {}    // rest of the '+' line
        if (c == -1) return -2;
        scratch_.clear();
        while (until(kLine, scratch_, nullptr, true) >= 0 && scratch_.size() < seq.size()) {}
        last_char_ = 0;
        if (scratch_.size() != seq.size()) return -2;
        return (int) seq.size();
    }

  private:
    enum { kSpace, kLine };

    bool fill() {
        if (eof_) return false;
        int n = gzread(fp_, buf_.data(), (unsigned) buf_.size());
        begin_ = 0;
        if (n <= 0) { eof_ = true; end_ = 0; err_ = n < 0; return false; }
        end_ = (size_t) n;
        return true;
    }
    int getc() {
        if (err_) return -3;
        if (begin_ >= end_ && !fill()) return err_ ? -3 : -1;
        return (unsigned char) buf_[begin_++];
    }
    // ks_getuntil2: read up to a delimiter (not stored); a trailing '\r' of a line is dropped.
    int until(int mode, std::string& str, int* dret, bool append) {
        bool gotany = false;
        if (dret) *dret = 0;
        if (!append) str.clear();
        for (;;) {
            if (err_) return -3;
            if (begin_ >= end_ && !fill()) { if (err_) return -3; break; }
            size_t i = begin_;
            if (mode == kLine) {
                const void* p = memchr(buf_.data() + begin_, '\n', end_ - begin_);
                i = p ? (size_t) ((const char*) p - buf_.data()) : end_;
            } else {
                while (i < end_ && !isspace((unsigned char) buf_[i])) ++i;
            }
            gotany = true;
            str.append(buf_.data() + begin_, i - begin_);
            begin_ = i + 1;
            if (i < end_) { if (dret) *dret = (unsigned char) buf_[i]; break; }
        }
        if (!gotany && eof_ && begin_ >= end_) return -1;
        if (mode == kLine && str.size() > 1 && str.back() == '\r') str.pop_back();
        return (int) str.size();
    }

    gzFile fp_;
    std::vector<char> buf_;
    std::string scratch_;
    size_t begin_ = 0, end_ = 0;
    bool eof_ = false, err_ = false;
    int last_char_ = 0;
};

// DocList (include/doclist.hpp:46-73): `name start` pairs; a text offset resolves to the
// document with the greatest start <= offset.
class DocList {
  public:
    bool load(const std::string& path) {
        std::ifstream ifs(path);
        if (!ifs.good()) return false;
        std::string name;
        uint64_t pos = 0;
        while (ifs >> name >> pos) { names_.push_back(name); starts_.push_back(pos); }
        std::sort(starts_.begin(), starts_.end());
        starts_.erase(std::unique(starts_.begin(), starts_.end()), starts_.end());
        return true;
    }
    // doc_and_offset_at; requires at least one document starting at or before i
    void resolve(uint64_t i, const std::string*& name, uint64_t& off) const {
        size_t rank = (size_t) (std::upper_bound(starts_.begin(), starts_.end(), i) - starts_.begin());
        if (rank == 0) rank = 1;
        name = &names_[std::min(rank, names_.size()) - 1];
        off = i - starts_[rank - 1];
    }
    bool empty() const { return names_.empty(); }

  private:
    std::vector<std::string> names_;
    std::vector<uint64_t> starts_;
};

// The whole buffer to fd with write(2): the report is produced in MB-sized slices, stdio would only copy them again.
inline void write_all(int fd, const char* p, size_t n) {
    while (n) {
        const ssize_t w = ::write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; perror("write"); exit(1); }
        p += w;
        n -= (size_t) w;
    }
}

inline void put_u64(std::string& out, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
    while (n) out.push_back(tmp[--n]);
}

template <class T>
class Channel {     // bounded MPMC queue
  public:
    explicit Channel(size_t cap) : cap_(cap) {}
    void push(T v) {
        std::unique_lock<std::mutex> l(m_);
        not_full_.wait(l, [&] { return q_.size() < cap_; });
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    bool pop(T& v) {
        std::unique_lock<std::mutex> l(m_);
        not_empty_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void close() {
        std::lock_guard<std::mutex> l(m_);
        closed_ = true;
        not_empty_.notify_all();
    }

  private:
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

}  // namespace rbhost
