// Host-side I/O of the rb_align driver: FASTA/FASTQ(.gz) record reader with klib-kseq
// semantics (what the reference's loop sees, include/kseq.h:178-218), the document list
// (include/doclist.hpp) and allocation-free number formatting for the stdout grammar of
// src/rb_align.cpp:118-145.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cerrno>
#include <zlib.h>

#include <map>
#include <memory>
#include <thread>

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

// BGZF input (bgzip: a gzip file made of independent <= 64 KB members, each carrying its compressed size in a
// 'BC' extra field).  A plain .gz is one deflate stream and inflates on one core (0.2 GB/s of FASTQ, the bound of
// the whole driver); BGZF blocks are inflated here by several threads, in groups, and handed out in order through
// the same read() contract as gzread.  Anything that is not a well-formed BGZF block ends the stream with an error.
class BgzfSource {
  public:
    static bool is_bgzf(const char* path) {
        unsigned char h[18];
        FILE* f = fopen(path, "rb");
        if (!f) return false;
        const size_t got = fread(h, 1, sizeof h, f);
        fclose(f);
        return got == sizeof h && block_size(h, sizeof h) > 0;
    }
    // start_upos: byte of the INFLATED stream the first read() starts at (FastxReader::seek: the block headers are walked
    // up to the block that holds it; a malformed header on the way is left to the workers to report)
    BgzfSource(const char* path, int threads, uint64_t start_upos = 0) {
        fd_ = ::open(path, O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0 || st.st_size == 0) { failed_ = true; return; }
        size_ = (size_t) st.st_size;
        void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) { failed_ = true; return; }
        data_ = (const unsigned char*) m;
        madvise(m, size_, MADV_SEQUENTIAL);
        for (uint64_t upos = 0; start_upos && scan_ < size_;) {
            const size_t bs = block_size(data_ + scan_, size_ - scan_);
            if (bs < 26 || scan_ + bs > size_) break;
            uint32_t isize;
            memcpy(&isize, data_ + scan_ + bs - 4, 4);
            if (upos + isize > start_upos) { skip_ = (size_t) (start_upos - upos); break; }
            upos += isize;
            scan_ += bs;
        }
        const int n = std::max(1, std::min(threads, 16));
        max_ahead_ = 4 * (size_t) n;
        for (int t = 0; t < n; ++t) workers_.emplace_back([this] { work(); });
    }
    ~BgzfSource() {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        cv_done_.notify_all();
        for (auto& w : workers_) w.join();
        if (data_) munmap((void*) data_, size_);
        if (fd_ >= 0) ::close(fd_);
    }
    bool ok() const { return !failed_; }
    // gzread's contract: bytes copied (0 at the end of the stream), -1 on a malformed stream
    int read(char* dst, unsigned want) {
        unsigned got = 0;
        while (got < want) {
            if (!cur_ || cur_pos_ >= cur_->out.size()) {
                if (cur_) { std::lock_guard<std::mutex> l(m_); tasks_.erase(next_take_ - 1); cur_ = nullptr; cv_work_.notify_all(); }
                std::unique_lock<std::mutex> l(m_);
                cv_done_.wait(l, [&] {
                    auto it = tasks_.find(next_take_);
                    return stop_ || (it != tasks_.end() && it->second->done) || (all_issued_ && next_take_ >= next_issue_);
                });
                auto it = tasks_.find(next_take_);
                if (it == tasks_.end() || !it->second->done) break;         // end of the stream
                if (it->second->bad) {                                    // deliver what was read before the bad block first
                    if (got) return (int) got;
                    return -1;
                }
                cur_ = it->second.get();
                cur_pos_ = std::min(skip_, cur_->out.size());             // only the first task of a stream opened at a position
                skip_ = 0;
                ++next_take_;
            }
            const size_t n = std::min<size_t>(want - got, cur_->out.size() - cur_pos_);
            if (n) memcpy(dst + got, cur_->out.data() + cur_pos_, n);        // a group of empty blocks has no buffer at all
            cur_pos_ += n;
            got += (unsigned) n;
        }
        return (int) got;
    }

    // Block table of a file made of well-formed BGZF blocks and nothing else: coff[i] / uoff[i] = where block i starts in
    // the file / in the inflated stream, one more entry for the ends.  false: some header is malformed or bytes trail
    // the last block (such files are read through read(), which reports the error where kseq would).
    static bool index(const unsigned char* data, size_t size, std::vector<uint64_t>& coff, std::vector<uint64_t>& uoff) {
        coff.clear();
        uoff.clear();
        uint64_t at = 0, u = 0;
        while (at < size) {
            const size_t bs = block_size(data + at, size - at);
            if (bs < 26 || at + bs > size) return false;
            const size_t xlen = data[at + 10] | ((size_t) data[at + 11] << 8);
            uint32_t isize;
            memcpy(&isize, data + at + bs - 4, 4);
            if (bs < 12 + xlen + 8 + 2 || isize > (1u << 16)) return false;
            coff.push_back(at);
            uoff.push_back(u);
            at += bs;
            u += isize;
        }
        coff.push_back(at);
        uoff.push_back(u);
        return true;
    }
    // Inflates the block at h (bs bytes, as block_size returned them) into out[0, isize); false: malformed / CRC mismatch.
    static bool inflate_block(const unsigned char* h, size_t bs, char* out, uint32_t isize) {
        const size_t xlen = h[10] | ((size_t) h[11] << 8);
        if (bs < 12 + xlen + 8 + 2) return false;                // BSIZE smaller than its own header + trailer: crafted block
        const unsigned char* cdata = h + 12 + xlen;
        const size_t clen = bs - 12 - xlen - 8;
        uint32_t crc, want;
        memcpy(&crc, h + bs - 8, 4);
        memcpy(&want, h + bs - 4, 4);
        if (want != isize) return false;
        z_stream z;
        memset(&z, 0, sizeof z);
        if (inflateInit2(&z, -15) != Z_OK) return false;
        z.next_in = const_cast<unsigned char*>(cdata);
        z.avail_in = (uInt) clen;
        unsigned char none[8];                                   // the end-of-file block inflates to nothing
        z.next_out = isize ? (unsigned char*) out : none;
        z.avail_out = isize ? isize : (uInt) sizeof none;
        const int rc = inflate(&z, Z_FINISH);
        const bool fine = rc == Z_STREAM_END && z.total_out == isize;
        inflateEnd(&z);
        return fine && crc32(crc32(0L, Z_NULL, 0), (const unsigned char*) out, isize) == crc;
    }
    // total size of the BGZF block whose header starts at h (0: not a BGZF block)
    static size_t block_size(const unsigned char* h, size_t avail) {
        if (avail < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
        const size_t xlen = h[10] | ((size_t) h[11] << 8);
        if (avail < 12 + xlen) return 0;
        for (size_t p = 12; p + 4 <= 12 + xlen;) {
            const size_t slen = h[p + 2] | ((size_t) h[p + 3] << 8);
            if (h[p] == 'B' && h[p + 1] == 'C' && slen == 2 && p + 6 <= 12 + xlen) return (size_t) (h[p + 4] | ((size_t) h[p + 5] << 8)) + 1;
            p += 4 + slen;
        }
        return 0;
    }

  private:
    struct Task {
        size_t begin = 0, end = 0;          // compressed byte range: whole blocks
        std::vector<char> out;
        bool done = false, bad = false;
    };
    // next group of blocks (about 2 MB of output); called with m_ held
    bool issue(std::shared_ptr<Task>& t) {
        if (all_issued_) return false;
        t = std::make_shared<Task>();
        t->begin = scan_;
        size_t blocks = 0;
        while (scan_ < size_ && blocks < 32) {
            const size_t bs = block_size(data_ + scan_, size_ - scan_);
            if (bs < 26 || scan_ + bs > size_) { t->bad = blocks == 0; all_issued_ = true; trailing_garbage_ = blocks != 0; break; }
            scan_ += bs;
            ++blocks;
        }
        if (scan_ >= size_) all_issued_ = true;
        t->end = scan_;
        if (blocks == 0 && !t->bad) return false;
        tasks_[next_issue_++] = t;
        if (trailing_garbage_) {                                   // a bad block after good ones: its own failing task
            auto bad = std::make_shared<Task>();
            bad->bad = bad->done = true;
            tasks_[next_issue_++] = bad;
        }
        return true;
    }
    void inflate_task(Task& t) {
        if (t.bad) return;
        size_t p = t.begin;
        while (p < t.end) {
            const unsigned char* h = data_ + p;
            const size_t bs = block_size(h, t.end - p);
            if (bs < 26) { t.bad = true; return; }
            uint32_t isize;
            memcpy(&isize, h + bs - 4, 4);
            if (isize > (1u << 16)) { t.bad = true; return; }
            const size_t at = t.out.size();
            t.out.resize(at + isize);
            if (!inflate_block(h, bs, t.out.data() + at, isize)) { t.bad = true; return; }
            p += bs;
        }
    }
    void work() {
        for (;;) {
            std::shared_ptr<Task> t;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_work_.wait(l, [&] { return stop_ || (!all_issued_ && tasks_.size() < max_ahead_); });
                if (stop_) return;
                if (!issue(t)) { cv_done_.notify_all(); if (all_issued_) return; continue; }
            }
            inflate_task(*t);
            {
                std::lock_guard<std::mutex> l(m_);
                t->done = true;
            }
            cv_done_.notify_all();
            if (all_issued_) { cv_done_.notify_all(); }
        }
    }

    int fd_ = -1;
    const unsigned char* data_ = nullptr;
    size_t size_ = 0, scan_ = 0;
    bool failed_ = false, stop_ = false, all_issued_ = false, trailing_garbage_ = false;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::map<uint64_t, std::shared_ptr<Task>> tasks_;
    uint64_t next_issue_ = 0, next_take_ = 0;
    size_t max_ahead_ = 8;
    Task* cur_ = nullptr;
    size_t cur_pos_ = 0, skip_ = 0;
    std::vector<std::thread> workers_;
};

// A plain .gz is one deflate stream: it inflates on one core, but it need not be the core that parses.  This runs
// gzread on its own thread, a few 1 MB blocks ahead of the record reader.
class GzAhead {
  public:
    explicit GzAhead(gzFile fp) : fp_(fp), th_([this] { run(); }) {}
    ~GzAhead() {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    // gzread's contract
    int read(char* dst, unsigned want) {
        unsigned got = 0;
        while (got < want) {
            if (pos_ >= cur_.size()) {
                std::unique_lock<std::mutex> l(m_);
                if (!cur_.empty()) { spare_.push_back(std::move(cur_)); cur_.clear(); cv_.notify_all(); }
                pos_ = 0;
                cv_.wait(l, [&] { return !ready_.empty() || done_; });
                if (ready_.empty()) return error_ && got == 0 ? -1 : (int) got;
                cur_ = std::move(ready_.front());
                ready_.pop_front();
                cv_.notify_all();
            }
            const size_t n = std::min<size_t>(want - got, cur_.size() - pos_);
            memcpy(dst + got, cur_.data() + pos_, n);
            pos_ += n;
            got += (unsigned) n;
        }
        return (int) got;
    }

  private:
    void run() {
        for (;;) {
            std::vector<char> buf;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || ready_.size() < 8; });
                if (stop_) return;
                if (!spare_.empty()) { buf = std::move(spare_.back()); spare_.pop_back(); }
            }
            buf.resize(1 << 20);
            const int n = gzread(fp_, buf.data(), (unsigned) buf.size());
            std::lock_guard<std::mutex> l(m_);
            if (n <= 0) {
                error_ = n < 0;
                done_ = true;
                cv_.notify_all();
                return;
            }
            buf.resize((size_t) n);
            ready_.push_back(std::move(buf));
            cv_.notify_all();
        }
    }
    gzFile fp_ = nullptr;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::vector<char>> ready_;
    std::vector<std::vector<char>> spare_;
    std::vector<char> cur_;
    size_t pos_ = 0;
    bool stop_ = false, done_ = false, error_ = false;
    std::thread th_;                    // last: starts when everything above exists
};

// Record reader.  Return codes of next() follow kseq_read: >=0 sequence length, -1 end of
// file, -2 truncated quality string, -3 stream error.
class FastxReader {
  public:
    // inflate_threads: 0 = all cores, 1 = plain gzread only (no BGZF workers, no read-ahead thread).
    // A FIFO, /dev/stdin or <(zcat x.fq.gz) is opened exactly ONCE, through gzopen as the reference does
    // (src/rb_align.cpp:169): every extra open/read of such a path would eat the first bytes of the stream, so the
    // gzip / BGZF probes below only touch regular files.
    explicit FastxReader(const char* path, int inflate_threads = 0) : path_(path), buf_(1 << 20) {
        struct stat st;
        const bool regular = stat(path, &st) == 0 && S_ISREG(st.st_mode);
        fp_ = gzopen(path, "r");
        if (fp_) gzbuffer(fp_, 1 << 18);
        if (inflate_threads <= 0) inflate_threads = (int) std::max(1u, std::thread::hardware_concurrency());
        inflate_threads_ = inflate_threads;
        if (!fp_ || !regular) return;
        if (inflate_threads > 1 && BgzfSource::is_bgzf(path)) {
            bgzf_.reset(new BgzfSource(path, inflate_threads));
            if (!bgzf_->ok()) bgzf_.reset();                     // fall back to the single zlib stream
        }
        unsigned char magic[2] = {0, 0};
        if (FILE* f = fopen(path, "rb")) {
            if (fread(magic, 1, 2, f) != 2) magic[0] = 0;
            fclose(f);
        }
        compressed_ = inflate_threads > 1 && magic[0] == 0x1f && magic[1] == 0x8b;
    }
    ~FastxReader() {
        ahead_.reset();                         // join the inflate thread before its gzFile goes away
        if (fp_) gzclose(fp_);
    }
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;
    bool ok() const { return fp_ != nullptr; }
    // Restart between records at byte `pos` of the (uncompressed) stream.
    bool seek(size_t pos) {
        if (bgzf_) {                                               // BGZF: restart the block inflaters at the block holding pos
            bgzf_.reset();
            bgzf_.reset(new BgzfSource(path_.c_str(), inflate_threads_, pos));
            if (!bgzf_->ok()) return false;
        } else {
            ahead_.reset();
            if (!fp_ || gzseek(fp_, (z_off_t) pos, SEEK_SET) < 0) return false;
        }
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
        if (compressed_ && !bgzf_ && !ahead_) ahead_.reset(new GzAhead(fp_));       // inflate on its own thread from here on
        int n = bgzf_ ? bgzf_->read(buf_.data(), (unsigned) buf_.size())
                      : ahead_ ? ahead_->read(buf_.data(), (unsigned) buf_.size()) : gzread(fp_, buf_.data(), (unsigned) buf_.size());
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

    std::string path_;
    int inflate_threads_ = 1;
    gzFile fp_ = nullptr;
    std::unique_ptr<BgzfSource> bgzf_;
    std::unique_ptr<GzAhead> ahead_;          // declared after fp_: destroyed (thread joined) before gzclose
    bool compressed_ = false;
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
    const std::vector<std::string>& names() const { return names_; }
    const std::vector<uint64_t>& starts() const { return starts_; }
    void set(std::vector<std::string> names, std::vector<uint64_t> starts) { names_ = std::move(names); starts_ = std::move(starts); }   // tests

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
    int n = 24;
    do { tmp[--n] = (char) ('0' + v % 10); v /= 10; } while (v);
    out.append(tmp + n, (size_t) (24 - n));
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
