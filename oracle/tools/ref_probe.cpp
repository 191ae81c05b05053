// Test infrastructure (NOT product code): calls the UNMODIFIED reference
// classes directly so that the oracle restatement and the CUDA path can be
// compared against reference *internals* (not just rb_align's stdout).
//
//   ref_probe runs     <prefix>   -> "n R" then one "head len" line per BWT run        (rle_string.hpp:224-242)
//   ref_probe f        <prefix>   -> the 256 entries of RowBowt's F array               (rowbowt.hpp:770-778)
//   ref_probe tsa      <prefix>   -> "r n last_run_sample" then "samples_last[j]" lines (toehold_sa.hpp:93-99)
//   ref_probe phi      <prefix>   < i per line      -> phi(i)                           (toehold_sa.hpp:56-72)
//   ref_probe range    <prefix>   < query per line  -> "lo hi"                          (rowbowt.hpp:121-131)
//   ref_probe toehold  <prefix>   < query per line  -> "lo hi k"                        (rowbowt.hpp:169-184)
//   ref_probe at_range <prefix>   < "s e" per line  -> marker words                     (rle_window_array.hpp:130-154)
//   ref_probe at       <prefix>   < "i" per line    -> marker words                     (rle_window_array.hpp:114-120)
//   ref_probe rank     <prefix>   < "i c" per line  -> rank(i,c) (c = byte value)       (rle_string.hpp:131-161)
#include <cstdio>
#include <iostream>
#include <fstream>
#include <string>
#include "rowbowt.hpp"
#include "rowbowt_io.hpp"
#include "rle_string.hpp"

using namespace std;

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: ref_probe <cmd> <prefix>\n"); return 2; }
    string cmd = argv[1], prefix = argv[2];
    ios::sync_with_stdio(false);
    if (cmd == "runs" || cmd == "rank") {
        ri::rle_string_sd bwt;
        ifstream ifs(prefix + ".rbwt", ios::binary);
        if (!ifs.good()) { cerr << "bad file\n"; return 1; }
        bwt.load(ifs);
        if (cmd == "runs") {
            cout << bwt.size() << " " << bwt.number_of_runs() << "\n";
            for (uint64_t j = 0; j < bwt.number_of_runs(); ++j) {
                auto rr = bwt.run_range(j);
                cout << (int) bwt[rr.first] << " " << (rr.second - rr.first + 1) << "\n";
            }
        } else {
            uint64_t i; int c;
            while (cin >> i >> c) cout << bwt.rank(i, (uint8_t) c) << "\n";
        }
        return 0;
    }
    if (cmd == "at_range" || cmd == "at") {
        MarkerArray<> ma;
        ifstream ifs(prefix + ".mab", ios::binary);
        if (!ifs.good()) { cerr << "bad file\n"; return 1; }
        ma.load(ifs);
        uint64_t s, e;
        vector<uint64_t> vals;
        if (cmd == "at_range") {
            while (cin >> s >> e) {
                vals.clear();
                ma.at_range(s, e, vals);
                for (auto v : vals) cout << v << " ";
                cout << "\n";
            }
        } else {
            while (cin >> s) {
                vals.clear();
                ma.at(s, vals);
                for (auto v : vals) cout << v << " ";
                cout << "\n";
            }
        }
        return 0;
    }
    if (cmd == "tsa" || cmd == "phi") {
        ToeholdSA tsa;
        ifstream ifs(prefix + ".tsa", ios::binary);
        if (!ifs.good()) { cerr << "bad file\n"; return 1; }
        tsa.load(ifs);
        if (cmd == "phi") {
            uint64_t i;
            while (cin >> i) cout << tsa.phi(i) << "\n";
        } else {
            uint64_t r, n;
            ifstream h(prefix + ".tsa", ios::binary);
            h.read((char*) &r, 8); h.read((char*) &n, 8);
            cout << r << " " << n << " " << tsa.get_last_run_sample() << "\n";
            for (uint64_t j = 0; j < r; ++j) cout << tsa.samples_last_at(j) << "\n";
        }
        return 0;
    }
    if (cmd == "f" || cmd == "range" || cmd == "toehold") {
        auto flag = cmd == "toehold" ? rbwt::LoadRbwtFlag::SA : rbwt::LoadRbwtFlag::NONE;
        rbwt::RowBowt<ri::rle_string_sd> rb(rbwt::load_rowbowt<ri::rle_string_sd>(prefix, flag));
        string q;
        if (cmd == "f") {
            // F[c] = first row of symbol c: recovered through LF on single symbols
            for (int c = 0; c < 256; ++c) {
                auto r = rb.LF(rb.full_range(), (uint8_t) c);
                cout << c << " " << r.first << " " << r.second << "\n";
            }
        } else if (cmd == "range") {
            while (getline(cin, q)) { auto r = rb.find_range(q); cout << r.first << " " << r.second << "\n"; }
        } else {
            while (getline(cin, q)) {
                auto lf = rb.find_range_w_toehold(q);
                cout << lf.rn.first << " " << lf.rn.second << " " << lf.ssamp << "\n";
            }
        }
        return 0;
    }
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}
