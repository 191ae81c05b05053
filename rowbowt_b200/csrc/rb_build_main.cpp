// rb_build — drop-in index builder over librowbowt_gpu (C ABI).  Same command line, inputs
// (<prefix>.bwt, .ssa/.esa with -s, .ma with -m, .docs with -l) and outputs (<out>.rbwt, .tsa, .mab,
// .docs, .ftab) as the reference builder (src/rb_build.cpp, rbwt::construct_and_serialize_rowbowt,
// include/rowbowt_io.hpp:49-89), and the files are byte-identical to the reference's.  The BWT is
// run-length encoded and the run-start samples are sorted on the GPU; the sdsl serialization is
// restated in sdsl_writer.hpp.
//
//   rb_build [-o out_prefix] [-s] [-m] [-l] [-f [-k K]] [--ftab-only] [--device D] <input_prefix>
//
// Not offered: --fbb (wt_fbb strings; the GPU path serves rle_string indexes only).
#include <getopt.h>
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "../../include/rowbowt_gpu.h"

namespace {

struct Args {                  // rbwt::RowBowtConstructArgs, include/rowbowt_io.hpp:33-47
    std::string inpre, prefix;
    int ma = 0, tsa = 0, dl = 0, ft = 0, ft_only = 0, fbb = 0, device = 0;
    size_t k = 10;
};

void print_help() {
    fprintf(stderr, "rb_build\n");
    fprintf(stderr, "Usage: rb_build [options] <index_prefix>\n");
    fprintf(stderr, "    --output_prefix/-o <basename>    output prefix\n");
    fprintf(stderr, "    --tsa/-s <basename>                 build toehold suffix array\n");
    fprintf(stderr, "    --ma/-m <basename>                  build marker array\n");
    fprintf(stderr, "    --ftab/-f <basename>                construct offset table (for faster querying)\n");
    fprintf(stderr, "    --device <D>                        CUDA device to build on (default 0)\n");
    fprintf(stderr, "    <input_prefix>                   index prefix\n");
}

bool file_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

[[noreturn]] void die_rbg(const char* what) {
    fprintf(stderr, "%s: %s\n", what, rbg_last_error());
    exit(1);
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    static struct option lopts[] = {{"output-prefix", required_argument, 0, 'o'},
                                    {"tsa", no_argument, 0, 's'},
                                    {"dl", no_argument, 0, 'l'},
                                    {"ftab-only", no_argument, 0, 'a'},
                                    {"ma", no_argument, 0, 'm'},
                                    {"ft", no_argument, 0, 'f'},
                                    {"fbb", no_argument, 0, 'x'},
                                    {"device", required_argument, 0, 'd'},
                                    {0, 0, 0, 0}};
    int c, li = 0;
    while ((c = getopt_long(argc, argv, "xo:k:lfsmha", lopts, &li)) != -1) {
        switch (c) {
            case 'x': a.fbb = 1; break;
            case 'o': a.prefix = optarg; break;
            case 's': a.tsa = 1; break;
            case 'm': a.ma = 1; break;
            case 'l': a.dl = 1; break;
            case 'f': a.ft = 1; break;
            case 'a': a.ft_only = 1; break;
            case 'k': a.k = std::stoull(optarg); break;
            case 'd': a.device = atoi(optarg); break;
            case 'h': print_help(); exit(0);
            case '?': break;
            default: print_help(); exit(1);
        }
    }
    if (argc - optind < 1) {
        fprintf(stderr, "no argument provided\n");
        exit(1);
    }
    a.inpre = argv[optind++];
    if (a.prefix.empty()) a.prefix = a.inpre;
    if (a.fbb) {
        fprintf(stderr, "--fbb (wt_fbb strings) is not supported by the GPU builder\n");
        return 1;
    }
    if (rbg_device_count() <= 0) {
        fprintf(stderr, "no CUDA device available (this build has no CPU path)\n");
        return 1;
    }
    if (a.ft_only) {           // construct_and_serialize_ftab, include/rowbowt_io.hpp:127-144
        rbg_index* ix = nullptr;
        if (file_exists(a.prefix + ".rbwt")) {
            std::cerr << "loading rbwt file" << std::endl;
            if (rbg_index_open(a.prefix.c_str(), RBG_LOAD_NONE, a.device, &ix) != RBG_OK) die_rbg("rbg_index_open");
        } else if (rbg_index_open_raw(a.inpre.c_str(), RBG_LOAD_NONE, a.device, &ix) != RBG_OK) {
            die_rbg("rbg_index_open_raw");
        }
        if (rbg_ftab_build(ix, (uint32_t) a.k) != RBG_OK) die_rbg("rbg_ftab_build");
        if (rbg_ftab_save(ix, (a.prefix + ".ftab").c_str()) != RBG_OK) die_rbg("rbg_ftab_save");
        rbg_index_close(ix);
        return 0;
    }
    std::cerr << "constructing using rle_string (GPU run-length encoder)" << std::endl;
    // the reference exits on a missing part (file_ne_error, include/rowbowt_io.hpp:28-31)
    auto need = [](const std::string& f) {
        if (!file_exists(f)) {
            std::cerr << "file " << f << " does not exist!" << std::endl;
            exit(1);
        }
    };
    if (a.ma) need(a.inpre + ".ma");
    if (a.tsa) { need(a.inpre + ".ssa"); need(a.inpre + ".esa"); }
    const uint32_t flags = (a.tsa ? RBG_LOAD_SA : 0) | (a.ma ? RBG_LOAD_MA : 0) | (a.ft ? RBG_LOAD_FT : 0);
    rbg_build_stats st;
    if (rbg_build_index(a.inpre.c_str(), a.prefix.c_str(), flags, (uint32_t) a.k, a.device, &st) != RBG_OK) die_rbg("rbg_build_index");
    if (a.dl) {                // rowbowt_io.hpp:73-81
        const std::string src = a.inpre + ".docs", dst = a.prefix + ".docs";
        if (src != dst) {
            std::ifstream ifs(src);
            std::ofstream ofs(dst);
            ofs << ifs.rdbuf();
        }
    }
    fprintf(stderr, "n=%llu r=%llu  rle %.3fs (read %.3fs, H2D+kernels %.1f ms)  samples %.3fs  markers %.3fs  write %.3fs  total %.3fs\n",
            (unsigned long long) st.n, (unsigned long long) st.r, st.s_rle, st.s_bwt_read, st.ms_rle_kernels, st.s_samples,
            st.s_markers, st.s_write, st.s_total);
    return 0;
}
