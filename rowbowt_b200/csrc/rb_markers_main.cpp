// rb_markers — drop-in host driver of the greedy-seeding marker genotyping over librowbowt_gpu
// (C ABI).  Same command line, index files (.rbwt/.mab[/.ftab]) and stdout grammar as the
// reference driver (src/rb_markers.cpp); the per-read worker body (:375-404) runs on the GPU a
// batch at a time through rbg_markers_greedy, both strands of every read.
//
// Output order: the reference prints each worker thread's buffer whenever it exceeds 4 KB, so its
// line order depends on thread scheduling; this driver always prints reads in input order, which
// is the reference's order at --threads 1.
//
// Not offered: --lmem (experimental O(m^2) path that logs every step to stderr), --overlap (the
// reference itself exits with "overlapped seeds currently broken").
#include <getopt.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include <atomic>
#include <condition_variable>
#include <mutex>

#include "../../include/rowbowt_gpu.h"
#include "fastx_parallel.hpp"
#include "host_io.hpp"

namespace {

struct Args {                       // RbAlignArgs, src/rb_markers.cpp:21-39
    std::string inpre, fastq;
    int ftab = 0, fbb = 0, overlap = 0, lmem = 0;
    int layout_cache = 0;           // keep / use <index_prefix>.rbgcache (RBG_LOAD_CACHE)
    size_t wsize = 19, max_range = 1000, min_range = 0, threads = 1, max_tasks = 1024, read_len = 101, min_seed_len = 0;
    int clear_conflicting = 0, clear_identical = 0, best_strand = 0, heuristic = 0;
    int gpus = 1;
    size_t batch_reads = 1u << 18;
};

void print_help() {
    fprintf(stderr, "rb_markers\n");
    fprintf(stderr, "Usage: rb_markers_only [options] <index_prefix> <input_fastq_name>\n");
    fprintf(stderr, "    --wsize            <int>         window size for performing marker queries along read\n");
    fprintf(stderr, "    --max-range        <int>         range-size upper threshold for performing marker queries\n");
    fprintf(stderr, "    --min-range        <int>         range-size upper threshold for performing marker queries\n");
    fprintf(stderr, "    --ftab                           seed through <index_prefix>.ftab\n");
    fprintf(stderr, "    --layout-cache                   keep the GPU layout in <index_prefix>.rbgcache and open from it when it is current\n");
    fprintf(stderr, "    --heuristic [--best-strand-only --min-seed-length <int> --clear-conflicting --clear-identical --read-len <int>]\n");
    fprintf(stderr, "    --gpus <N> --batch <reads>       GPUs to use (index replicated), reads per batch\n");
    fprintf(stderr, "    <input_prefix>                   index prefix\n");
    fprintf(stderr, "    <input_fastq>                    input fastq\n");
}

Args parse_args(int argc, char** argv) {
    Args a;
    static struct option lopts[] = {{"wsize", required_argument, 0, 'w'},
                                    {"max-range", required_argument, 0, 'r'},
                                    {"min-range", required_argument, 0, 'm'},
                                    {"threads", required_argument, 0, 't'},
                                    {"max-tasks", required_argument, 0, 'u'},
                                    {"read-len", required_argument, 0, 'l'},
                                    {"fbb", no_argument, &a.fbb, 1},
                                    {"ftab", no_argument, &a.ftab, 1},
                                    {"layout-cache", no_argument, &a.layout_cache, 1},
                                    {"overlap", no_argument, &a.overlap, 1},
                                    {"lmem", no_argument, &a.lmem, 1},
                                    {"heuristic", no_argument, &a.heuristic, 1},
                                    {"best-strand-only", no_argument, &a.best_strand, 1},
                                    {"min-seed-length", required_argument, 0, 'y'},
                                    {"clear-conflicting", no_argument, &a.clear_conflicting, 1},
                                    {"clear-identical", no_argument, &a.clear_identical, 1},
                                    {"gpus", required_argument, 0, 'g'},
                                    {"batch", required_argument, 0, 'b'},
                                    {0, 0, 0, 0}};
    int c, li = 0;
    while ((c = getopt_long(argc, argv, "o:w:r:hft:m:u:xl:y:g:b:", lopts, &li)) != -1) {
        switch (c) {
            case 0: break;
            case 'y': a.min_seed_len = std::atol(optarg); break;
            case 'l': a.read_len = std::atol(optarg); break;
            case 't': a.threads = std::atol(optarg); break;         // here: host parser + formatter threads (the GPU replaces the worker pool)
            case 'u': a.max_tasks = std::atol(optarg); break;
            case 'f': a.ftab = 1; break;
            case 'r': a.max_range = std::atol(optarg); break;
            case 'm': a.min_range = std::atol(optarg); break;
            case 'w': a.wsize = std::atol(optarg); break;
            case 'o': break;
            case 'h': print_help(); exit(0);
            case 'x': a.fbb = 1; break;
            case 'g': a.gpus = std::max(1, atoi(optarg)); break;
            case 'b': a.batch_reads = (size_t) std::max(1ll, atoll(optarg)); break;
            default: print_help(); exit(1);
        }
    }
    if (a.overlap) {                        // src/rb_markers.cpp:119-122
        fprintf(stderr, "overlapped seeds currently broken\n");
        exit(1);
    }
    if (a.lmem) {
        fprintf(stderr, "--lmem is not offered by the GPU driver\n");
        exit(1);
    }
    if (argc - optind < 2) {
        fprintf(stderr, "no argument provided\n");
        exit(1);
    }
    a.inpre = argv[optind++];
    a.fastq = argv[optind++];
    return a;
}

// MarkerT accessors, pfbwt-f/include/marker.hpp:7-37
inline uint64_t m_seq(uint64_t m) { return (m & 0x0FFFF00000000000ull) >> 46; }
inline uint64_t m_pos(uint64_t m) { return m & 0x00000FFFFFFFFFFFull; }
inline uint64_t m_allele(uint64_t m) { return (m & 0xF000000000000000ull) >> 60; }

using rbhost::ReadBatch;

struct Seed {                           // MarkerSeed, src/rb_markers.cpp:255-300
    int strand;                         // 0 = '+', 1 = '-'
    uint64_t range_size, query_start, query_len;
    std::vector<uint64_t> markers;
};

void print_seed(std::string& o, const char* name, size_t name_len, const Seed& s) {      // MarkerSeed::print_buf :262-272
    o.append(name, name_len);
    o += ' ';
    rbhost::put_u64(o, s.range_size);
    o += s.strand ? " - " : " + ";
    rbhost::put_u64(o, s.query_start);
    o += ' ';
    rbhost::put_u64(o, s.query_len);
    if (!s.markers.empty()) {
        for (uint64_t m : s.markers) {
            o += ' ';
            rbhost::put_u64(o, m_seq(m));
            o += '/';
            rbhost::put_u64(o, m_pos(m));
            o += '/';
            rbhost::put_u64(o, m_allele(m));
        }
    } else {
        o += " .";
    }
    o += '\n';
}

Seed make_seed(const rbg_seed_result& r, uint64_t j, int strand) {
    const rbg_seed& g = r.seeds[j];
    Seed s;
    s.strand = strand;
    s.range_size = g.hi - g.lo + 1;
    s.query_start = g.query_start == 0xFFFFFFFFu ? ~0ull : g.query_start;      // the reference's size_t(-1)
    s.query_len = g.query_len;
    s.markers.assign(r.markers + g.mk_off, r.markers + g.mk_off + g.mk_cnt);
    return s;
}

// MarkerSeed::filter_identical_pos, src/rb_markers.cpp:276-288 (markers sorted)
void filter_identical_pos(std::vector<uint64_t>& mk) {
    if (mk.empty()) return;
    uint64_t pm = 0;
    std::vector<uint64_t> keep;
    for (size_t i = 0; i < mk.size(); ++i) {
        const uint64_t m = mk[i];
        if (m_seq(m) == m_seq(pm) && m_pos(m) == m_pos(pm)) continue;
        pm = m;
        if (i + 1 < mk.size() && m_seq(mk[i + 1]) == m_seq(m) && m_pos(mk[i + 1]) == m_pos(m)) continue;
        keep.push_back(m);
    }
    mk.swap(keep);
}

// MarkerSeed::clear_if_conflicting, :291-296
void clear_if_conflicting(std::vector<uint64_t>& mk, size_t read_len) {
    if (mk.empty()) return;
    if (m_seq(mk.back()) != m_seq(mk.front()) || m_pos(mk.back()) - m_pos(mk.front()) >= read_len) mk.clear();
}

// One seed of the plain worker straight from the result arrays (no Seed object: this path prints ~10 seeds per read).
void print_raw_seed(std::string& o, const char* name, size_t name_len, const rbg_seed_result& r, uint64_t j, int strand) {
    const rbg_seed& g = r.seeds[j];
    o.append(name, name_len);
    o += ' ';
    rbhost::put_u64(o, g.hi - g.lo + 1);
    o += strand ? " - " : " + ";
    rbhost::put_u64(o, g.query_start == 0xFFFFFFFFu ? ~0ull : (uint64_t) g.query_start);     // the reference's size_t(-1)
    o += ' ';
    rbhost::put_u64(o, g.query_len);
    if (g.mk_cnt) {
        for (uint64_t k = 0; k < g.mk_cnt; ++k) {
            const uint64_t m = r.markers[g.mk_off + k];
            o += ' ';
            rbhost::put_u64(o, m_seq(m));
            o += '/';
            rbhost::put_u64(o, m_pos(m));
            o += '/';
            rbhost::put_u64(o, m_allele(m));
        }
    } else {
        o += " .";
    }
    o += '\n';
}

// The report of reads [i0, i1) of one batch.
void format_slice(const Args& a, const rbg_seed_result& r, const ReadBatch& b, uint64_t i0, uint64_t i1, std::string& o) {
    o.clear();
    o.reserve((size_t) ((r.seed_off[2 * i1] - r.seed_off[2 * i0]) * 40 + (i1 - i0) * 16));
    std::vector<Seed> seeds;
    for (uint64_t i = i0; i < i1; ++i) {
        size_t nl;
        const char* nm = b.name(i, nl);
        if (!a.heuristic) {                                     // worker, :347-415
            for (int s = 0; s < 2; ++s)
                for (uint64_t j = r.seed_off[2 * i + s]; j < r.seed_off[2 * i + s + 1]; ++j) print_raw_seed(o, nm, nl, r, j, s);
            continue;
        }
        // worker_heuristic, :416-507: a random strand first, the other one only if no seed of the first asked to stop
        seeds.clear();
        bool stop = false;
        const int first = b.aux[i] ? 1 : 0;
        for (int pass = 0; pass < 2 && !(pass == 1 && stop); ++pass) {
            const int s = pass == 0 ? first : 1 - first;
            for (uint64_t j = r.seed_off[2 * i + s]; j < r.seed_off[2 * i + s + 1]; ++j) {
                Seed ms = make_seed(r, j, s);
                if (ms.query_len < a.min_seed_len) continue;
                if (a.clear_conflicting) clear_if_conflicting(ms.markers, a.read_len);
                if (a.clear_identical) filter_identical_pos(ms.markers);
                const uint64_t used = ms.query_start + ms.query_len;
                seeds.push_back(std::move(ms));
                if (a.best_strand && (uint64_t) a.read_len - used < a.min_seed_len) stop = true;
            }
        }
        if (a.best_strand && !seeds.empty()) {                  // SeedVec::keep_seeds_best_strand, :306-325
            size_t best = 0;
            for (size_t k = 1; k < seeds.size(); ++k)
                if (seeds[best].query_len < seeds[k].query_len) best = k;
            const int strand = seeds[best].strand;
            std::vector<Seed> kept;
            for (auto& s : seeds)
                if (s.strand == strand) kept.push_back(std::move(s));
            seeds.swap(kept);
        }
        for (const Seed& s : seeds) print_seed(o, nm, nl, s);
    }
}

using rbhost::Channel;

}  // namespace

int main(int argc, char** argv) {
    Args args = parse_args(argc, argv);
    using clk = std::chrono::high_resolution_clock;
    auto t0 = clk::now();
    std::cerr << "(gpu) loading rowbowt + markers";
    if (args.ftab) std::cerr << " and ftab";
    std::cerr << std::endl;
    const int ndev = rbg_device_count();
    if (ndev <= 0) {
        fprintf(stderr, "no CUDA device available (this build has no CPU path)\n");
        return 1;
    }
    const int gpus = std::min(args.gpus, ndev);
    const uint32_t flags = RBG_LOAD_MA | (args.ftab ? RBG_LOAD_FT : 0) | (args.fbb ? RBG_LOAD_FBB : 0) | (args.layout_cache ? RBG_LOAD_CACHE : 0);       // load_rbwt, src/rb_markers.cpp:534-542
    std::vector<rbg_index*> idx(gpus, nullptr);
    for (int g = 0; g < gpus; ++g) {
        if (rbg_index_open(args.inpre.c_str(), flags, g, &idx[g]) != RBG_OK) {
            std::cerr << rbg_last_error() << std::endl;
            return 1;
        }
    }
    std::chrono::duration<double> diff = clk::now() - t0;
    std::cerr << "loading rowbowt + markers took: " << diff.count() << " seconds\n";
    t0 = clk::now();
    // host pipeline as in rb_align: parser threads over the mmap'ed FASTQ -> GPU workers -> formatter pool -> ordered writer
    const int host_threads = args.threads > 1 ? (int) args.threads : (int) std::max(1u, std::thread::hardware_concurrency());
    const size_t pool = 2 * (size_t) host_threads + 4 * (size_t) gpus + 4;
    rbhost::FastxBatchSource src(args.fastq.c_str(), host_threads, 0, args.batch_reads, rbhost::HostAlloc{rbg_host_alloc, rbg_host_free}, pool);
    if (!src.ok()) {
        fprintf(stderr, "invalid file\n");
        return 1;
    }
    rbg_greedy_params gp{args.wsize, args.max_range, args.min_range, (uint32_t) args.ftab, 0};

    struct Job {                                   // one batch between its query and its last formatted slice
        std::unique_ptr<ReadBatch> b;
        rbg_seed_result res;
        std::atomic<int> left{0};
        int gpu = 0;
    };
    struct Slice { Job* job; int s; uint64_t i0, i1; };
    Channel<std::unique_ptr<ReadBatch>> to_gpu(2 * gpus), to_writer(pool);
    Channel<Slice> to_format(64 * (size_t) host_threads);
    std::mutex inflight_m;
    std::condition_variable inflight_cv;
    std::vector<int> inflight(gpus, 0);            // seed results of device g not yet freed

    std::vector<std::thread> workers, formatters;
    for (int g = 0; g < gpus; ++g)
        workers.emplace_back([&, g] {
            std::unique_ptr<ReadBatch> b;
            while (to_gpu.pop(b)) {
                {
                    std::unique_lock<std::mutex> l(inflight_m);
                    inflight_cv.wait(l, [&] { return inflight[g] < 3; });
                    ++inflight[g];
                }
                Job* job = new Job;
                rbg_batch in{b->n, b->bases.p, b->offs.p};
                if (rbg_markers_greedy(idx[g], &in, &gp, &job->res) != RBG_OK) {
                    // the reference exits (k - 1 > wsize) or dies on an uncaught std::out_of_range (read shorter than k) here
                    fprintf(stderr, "ERROR: %s\n", rbg_last_error());
                    exit(1);
                }
                job->gpu = g;
                const uint64_t n = b->n;
                const int slices = (int) std::max<uint64_t>(1, std::min<uint64_t>((uint64_t) host_threads, n >> 11));
                b->out.resize(slices);
                job->b = std::move(b);
                job->left = slices;
                for (int s = 0; s < slices; ++s)
                    to_format.push(Slice{job, s, n * (uint64_t) s / slices, n * (uint64_t) (s + 1) / slices});
            }
        });
    for (int t = 0; t < host_threads; ++t)
        formatters.emplace_back([&] {
            Slice sl;
            while (to_format.pop(sl)) {
                Job* job = sl.job;
                format_slice(args, job->res, *job->b, sl.i0, sl.i1, job->b->out[sl.s]);
                if (job->left.fetch_sub(1) == 1) {
                    const int g = job->gpu;
                    rbg_seed_result_free(&job->res);
                    {
                        std::lock_guard<std::mutex> l(inflight_m);
                        --inflight[g];
                    }
                    inflight_cv.notify_all();
                    to_writer.push(std::move(job->b));
                    delete job;
                }
            }
        });
    std::thread writer([&] {
        std::map<uint64_t, std::unique_ptr<ReadBatch>> pending;
        uint64_t next = 0;
        std::unique_ptr<ReadBatch> b;
        while (to_writer.pop(b)) {
            pending[b->id] = std::move(b);
            for (auto it = pending.find(next); it != pending.end(); it = pending.find(next)) {
                for (const std::string& o : it->second->out) rbhost::write_all(1, o.data(), o.size());
                src.recycle(std::move(it->second));
                pending.erase(it);
                ++next;
            }
        }
    });

    // RandomBoolGenerator, src/rb_markers.cpp:225-240: one bit per read IN INPUT ORDER, 32 per draw of a default-seeded
    // mt19937 -- batches leave the source in input order, so the bits are dealt here, on the one thread that sees them all
    std::mt19937 rng;
    uint32_t bits = 0;
    int bit_count = 0;
    while (std::unique_ptr<ReadBatch> b = src.next()) {
        if (args.heuristic) {
            b->aux.resize(b->n);
            for (uint64_t i = 0; i < b->n; ++i) {
                if (bit_count == 0) { bits = (uint32_t) rng(); bit_count = 32; }
                b->aux[i] = (bits & 1u) ? 0 : 1;                // get_bool() ? FWD : REV
                bits >>= 1;
                --bit_count;
            }
        }
        to_gpu.push(std::move(b));
    }
    const int err = src.err();
    to_gpu.close();
    for (auto& w : workers) w.join();
    to_format.close();
    for (auto& f : formatters) f.join();
    to_writer.close();
    writer.join();
    for (auto* ix : idx) rbg_index_close(ix);
    switch (err) {                      // src/rb_markers.cpp:584-593
        case -2: fprintf(stderr, "ERROR: truncated quality string\n"); exit(1);
        case -3: fprintf(stderr, "ERROR: error reading stream\n"); exit(1);
        default: break;
    }
    diff = clk::now() - t0;
    std::cerr << "counting markers took: " << diff.count() << " seconds" << std::endl;
    return 0;
}
