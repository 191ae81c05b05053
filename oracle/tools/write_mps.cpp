// Test/bench infrastructure (NOT product code).
// Drives the reference's header-only MarkerPositionsWriter
// (/root/reference/pfbwt-f/include/marker_array.hpp:31-132) so that synthetic
// marker panels go through exactly the reference's .mps writer.
//
// stdin: text lines, either
//     <textpos> <refpos> <allele>      -> update(textpos, refpos, allele, 0)
//     -                                -> finish_sequence()
// usage: write_mps <wsize> <out.mps>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "marker_array.hpp"

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: write_mps <wsize> <out.mps> < positions.txt\n"); return 2; }
    size_t w = strtoull(argv[1], nullptr, 10);
    FILE* fp = fopen(argv[2], "wb");
    if (!fp) { perror("fopen"); return 1; }
    MarkerPositionsWriter writer(w, fp);
    char line[256];
    bool open_seq = false;
    while (fgets(line, sizeof line, stdin)) {
        if (line[0] == '-') { writer.finish_sequence(); open_seq = false; continue; }
        unsigned long long tp, rp; int gt;
        if (sscanf(line, "%llu %llu %d", &tp, &rp, &gt) != 3) continue;
        writer.update(tp, rp, gt, 0);
        open_seq = true;
    }
    if (open_seq) writer.finish_sequence();
    fclose(fp);
    return 0;
}
