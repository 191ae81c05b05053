// rb_align — drop-in host driver over librowbowt_gpu (C ABI).  Same command line, same
// index files (.rbwt/.tsa/.mab/.docs), same stdout grammar and stderr timing line as the
// reference driver (src/rb_align.cpp), but the per-read loop (src/rb_align.cpp:176-178) is
// batched: reads are parsed on the host, queried on the GPU(s) a batch at a time, and
// printed in FASTQ order.
//
//   rb_align [-s] [-m] [-o prefix] [--gpus N] [--batch READS] [--ftab | --ftab-k K] <index_prefix> <fastq>
//
// Pipeline: one parser thread (gz + kseq-compatible reader) -> N GPU workers (one index
// replica and one C-ABI handle per device; each also formats its batch's text) -> one
// ordered writer.  No collective: batches are independent.
#include <getopt.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rowbowt_gpu.h"
#include "host_io.hpp"

namespace {

struct Args {
    std::string inpre, fastq, outpre;
    int sam = 0, markers = 0, fbb = 0;
    int gpus = 1;
    size_t batch_reads = 1u << 20;
    int ftab_file = 0;          // load <prefix>.ftab (LoadRbwtFlag::FT) instead of building the seed table
    int ftab_k = 10;            // k of the seed table built on the GPU at load (RowBowt::build_ftab default); 0 = none
    int parse_only = 0;         // diagnostic: dump "name<TAB>sequence" per record, no GPU needed
};

void print_help() {
    fprintf(stderr, "rb_align");
    fprintf(stderr, "Usage: rb_align [options] <index_prefix> <input_fastq_name>\n");
    fprintf(stderr, "    --output_prefix/-o <basename>    output prefix\n");
    fprintf(stderr, "    --markers/-m                     print markers\n");
    fprintf(stderr, "    --sam/-s                         print locations\n");
    fprintf(stderr, "    --gpus/-g <N>                    number of GPUs (index replicated, batches sharded)\n");
    fprintf(stderr, "    --batch/-b <reads>               reads per GPU batch (default 1048576)\n");
    fprintf(stderr, "    --ftab                           load the k-mer seed table from <index_prefix>.ftab\n");
    fprintf(stderr, "    --ftab-k/-k <k>                  build the k-mer seed table on the GPU (default 10, 0 = none)\n");
    fprintf(stderr, "    <input_prefix>                   index prefix\n");
    fprintf(stderr, "    <input_fastq>                    input fastq\n");
}

Args parse_args(int argc, char** argv) {
    Args a;
    static struct option lopts[] = {{"output_prefix", required_argument, 0, 'o'},
                                    {"markers", no_argument, 0, 'm'},
                                    {"sam", no_argument, 0, 's'},
                                    {"fbb", no_argument, 0, 'f'},
                                    {"gpus", required_argument, 0, 'g'},
                                    {"batch", required_argument, 0, 'b'},
                                    {"parse-only", no_argument, 0, 'P'},
                                    {"ftab", no_argument, 0, 'F'},
                                    {"ftab-k", required_argument, 0, 'k'},
                                    {"help", no_argument, 0, 'h'},
                                    {0, 0, 0, 0}};
    int c, li = 0;
    while ((c = getopt_long(argc, argv, "o:smhg:b:k:", lopts, &li)) != -1) {
        switch (c) {
            case 'f': a.fbb = 1; break;
            case 'o': a.outpre = optarg; break;
            case 'h': print_help(); exit(0);
            case 's': a.sam = 1; break;
            case 'm': a.markers = 1; break;
            case 'P': a.parse_only = 1; break;
            case 'F': a.ftab_file = 1; break;
            case 'k': a.ftab_k = std::max(0, atoi(optarg)); break;
            case 'g': a.gpus = std::max(1, atoi(optarg)); break;
            case 'b': a.batch_reads = (size_t) std::max(1ll, atoll(optarg)); break;
            default: print_help(); exit(1);
        }
    }
    if (argc - optind < (a.parse_only ? 1 : 2)) {
        fprintf(stderr, "no argument provided\n");
        exit(1);
    }
    if (!a.parse_only) a.inpre = argv[optind++];
    a.fastq = argv[optind++];
    if (a.outpre.empty()) a.outpre = a.inpre;
    return a;
}

struct Batch {
    uint64_t id = 0;
    std::vector<std::string> names;
    std::string bases;
    std::vector<uint64_t> offs{0};
    std::string out;            // formatted text
};

// MarkerT accessors, pfbwt-f/include/marker.hpp:19-21,35-37
inline uint64_t marker_pos(uint64_t m) { return m & 0x00000FFFFFFFFFFFull; }
inline uint64_t marker_allele(uint64_t m) { return (m & 0xF000000000000000ull) >> 60; }

// rb_report's text for one batch, src/rb_align.cpp:118-145
void format_batch(const Args& args, const rbhost::DocList& docs, const rbg_result& r, Batch& b) {
    std::string& o = b.out;
    o.clear();
    o.reserve(b.names.size() * 48);
    for (size_t i = 0; i < b.names.size(); ++i) {
        o += b.names[i].c_str();                       // printed as a C string
        o += " (";
        rbhost::put_u64(o, r.lo[i]);
        o += ',';
        rbhost::put_u64(o, r.hi[i]);
        o += "), count=";
        rbhost::put_u64(o, r.hi[i] - r.lo[i] + 1);     // 64-bit wraparound, as printed by the reference
        o += '\n';
        if (args.sam) {
            o += "\tlocs: ";
            for (uint64_t j = r.loc_off[i]; j < r.loc_off[i + 1]; ++j) {
                const std::string* dn;
                uint64_t off;
                docs.resolve(r.locs[j], dn, off);
                rbhost::put_u64(o, r.locs[j]);
                o += '/';
                o += *dn;
                o += ':';
                rbhost::put_u64(o, off);
                o += ' ';
            }
            o += '\n';
        }
        if (args.markers) {
            o += "\tmarkers: ";
            if (r.mk_off[i] == r.mk_off[i + 1]) o += "no markers (consider building the marker array with a larger window size)";
            for (uint64_t j = r.mk_off[i]; j < r.mk_off[i + 1]; ++j) {
                rbhost::put_u64(o, marker_pos(r.markers[j]));
                o += '/';
                rbhost::put_u64(o, marker_allele(r.markers[j]));
                o += ' ';
            }
            o += '\n';
        }
    }
}

using rbhost::Channel;

[[noreturn]] void die_rbg(const char* what) {
    fprintf(stderr, "%s: %s\n", what, rbg_last_error());
    exit(1);
}

}  // namespace

int main(int argc, char** argv) {
    Args args = parse_args(argc, argv);
    if (args.parse_only) {
        rbhost::FastxReader rd(args.fastq.c_str());
        if (!rd.ok()) { fprintf(stderr, "invalid file\n"); return 1; }
        std::string name, seq;
        int err;
        while ((err = rd.next(name, seq)) >= 0) printf("%s\t%s\n", name.c_str(), seq.c_str());
        if (err == -2) { fprintf(stderr, "ERROR: truncated quality string\n"); return 1; }
        if (err == -3) { fprintf(stderr, "ERROR: error reading stream\n"); return 1; }
        return 0;
    }
    if (args.fbb) {
        fprintf(stderr, "--fbb indexes (wt_fbb .rbwt) are not supported by the GPU path\n");
        return 1;
    }
    using clk = std::chrono::high_resolution_clock;
    auto t0 = clk::now();

    // load_rbwt, src/rb_align.cpp:147-160
    uint32_t flags = 0;
    rbhost::DocList docs;
    if (args.sam) {
        std::cerr << "will load SA and DA" << std::endl;
        flags |= RBG_LOAD_SA | RBG_LOAD_DL;
    }
    if (args.markers) {
        std::cerr << "will load SA and DA" << std::endl;
        flags |= RBG_LOAD_MA;
    }
    if (args.ftab_file) flags |= RBG_LOAD_FT;
    int ndev = rbg_device_count();
    if (ndev <= 0) {
        fprintf(stderr, "no CUDA device available (this build has no CPU path)\n");
        return 1;
    }
    const int gpus = std::min(args.gpus, ndev);
    std::vector<rbg_index*> idx(gpus, nullptr);
    {
        std::vector<std::thread> th;
        std::vector<int> rc(gpus, 0);
        std::vector<std::string> msg(gpus);
        for (int g = 0; g < gpus; ++g)
            th.emplace_back([&, g] {
                rc[g] = rbg_index_open(args.inpre.c_str(), flags, g, &idx[g]);
                if (!rc[g] && !args.ftab_file && args.ftab_k) rc[g] = rbg_ftab_build(idx[g], (uint32_t) args.ftab_k);
                if (rc[g]) msg[g] = rbg_last_error();
            });
        for (auto& t : th) t.join();
        for (int g = 0; g < gpus; ++g)
            if (rc[g]) {
                // the reference prints "bad file" and exits for a missing part (rowbowt_io.hpp:166-169)
                std::cerr << (rc[g] == RBG_E_IO ? "bad file" : msg[g]) << std::endl;
                return 1;
            }
    }
    if (args.sam) {
        std::cerr << "loading: " << args.inpre + ".docs" << std::endl;
        if (!docs.load(args.inpre + ".docs")) {
            std::cerr << "bad file" << std::endl;
            return 1;
        }
    }
    std::chrono::duration<double> load_time = clk::now() - t0;

    rbhost::FastxReader reader(args.fastq.c_str());
    if (!reader.ok()) {
        fprintf(stderr, "invalid file\n");
        return 1;
    }
    auto q0 = clk::now();
    const uint32_t mode = (args.sam ? RBG_LOCATE : 0) | (args.markers ? RBG_MARKERS : 0);

    Channel<std::unique_ptr<Batch>> to_gpu(2 * gpus), to_writer(4 * gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < gpus; ++g)
        workers.emplace_back([&, g] {
            std::unique_ptr<Batch> b;
            while (to_gpu.pop(b)) {
                rbg_batch in{b->names.size(), b->bases.data(), b->offs.data()};
                rbg_result res;
                if (rbg_query(idx[g], &in, mode, UINT64_MAX, &res) != RBG_OK) die_rbg("rbg_query");
                format_batch(args, docs, res, *b);
                rbg_result_free(&res);
                to_writer.push(std::move(b));
            }
        });
    std::thread writer([&] {
        std::map<uint64_t, std::unique_ptr<Batch>> pending;
        uint64_t next = 0;
        std::unique_ptr<Batch> b;
        while (to_writer.pop(b)) {
            pending[b->id] = std::move(b);
            for (auto it = pending.find(next); it != pending.end(); it = pending.find(next)) {
                fwrite(it->second->out.data(), 1, it->second->out.size(), stdout);
                pending.erase(it);
                ++next;
            }
        }
        fflush(stdout);
    });

    int err;
    uint64_t bid = 0;
    std::unique_ptr<Batch> cur(new Batch);
    std::string name, seq;
    while ((err = reader.next(name, seq)) >= 0) {
        cur->names.push_back(name);
        cur->bases.append(seq.c_str());                // the reference passes seq.s as a C string
        cur->offs.push_back(cur->bases.size());
        if (cur->names.size() >= args.batch_reads) {
            cur->id = bid++;
            to_gpu.push(std::move(cur));
            cur.reset(new Batch);
        }
    }
    if (!cur->names.empty()) {
        cur->id = bid++;
        to_gpu.push(std::move(cur));
    }
    to_gpu.close();
    for (auto& w : workers) w.join();
    to_writer.close();
    writer.join();
    std::chrono::duration<double> query_time = clk::now() - q0;
    for (auto* ix : idx) rbg_index_close(ix);
    switch (err) {           // src/rb_align.cpp:181-191
        case -2: fprintf(stderr, "ERROR: truncated quality string\n"); exit(1);
        case -3: fprintf(stderr, "ERROR: error reading stream\n"); exit(1);
        default: break;
    }
    std::cerr << load_time.count() << " " << query_time.count() << std::endl;
    return 0;
}
