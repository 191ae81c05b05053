/*
 * ORACLE — test infrastructure, NOT product code.
 *
 * A plain-C, CPU restatement of the reference's (alshai/rowbowt) rb_align query
 * path over FLAT arrays.  It follows the reference's
 * algorithm function by function (file:line cited on each), with the succinct
 * containers (sdsl sd_vector / wt_huff) replaced by sorted arrays + binary
 * search that return the same values.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library; the product
 * (rowbowt_b200/) never does.
 *
 * Parity pin: tests/test_oracle.py checks every function here against the
 * reference's golden vectors (tests/rb_tests.cpp:47-58,115-120,131-140,147-173)
 * and against the compiled reference itself (oracle/_ref/rb_align, ref_probe).
 *
 * Paths are /root/reference-relative.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;

typedef struct {
    /* rle_string (include/rle_string.hpp:385-395) as flat runs */
    u64 n, R;
    uint8_t* head;      /* [R] */
    u64* start;         /* [R+1] start[j] = sum of len[0..j) */
    u64* crun[256];     /* crun[c][k]  = BWT-order index of the k-th c-run           (run_heads.select(k,c)) */
    u64* ccum[256];     /* ccum[c][k]  = total length of the first k c-runs, [nc+1]  (runs_per_letter[c])   */
    u64 nc[256];        /* number of c-runs */
    u64 F[257];         /* RowBowt::build_f, include/rowbowt.hpp:770-778 */
    /* ToeholdSA (include/toehold_sa.hpp:157-161) */
    u64 r;
    u64 *pred, *samples_last, *pred_to_run;
    int has_tsa;
    /* rle_window_arr (pfbwt-f/include/rle_window_array.hpp:258-264) */
    u64 sz_starts, sz_ends, sz_idxs;
    u64 nstarts, nends, nidxs;
    u64 *wstarts, *wends, *widxs, *arr;
    u64 arr_size;
    int has_ma;
    u64 lf_steps;       /* LF(range,c) calls executed since creation (stat only) */
} orc_index;

/* number of elements of sorted a[0..m) that are <  x */
static u64 lower_bound(const u64* a, u64 m, u64 x) {
    u64 lo = 0, hi = m;
    while (lo < hi) { u64 mid = lo + (hi - lo) / 2; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
/* number of elements of sorted a[0..m) that are <= x */
static u64 upper_bound(const u64* a, u64 m, u64 x) {
    u64 lo = 0, hi = m;
    while (lo < hi) { u64 mid = lo + (hi - lo) / 2; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
    return lo;
}

void orc_free(orc_index* ix) {
    if (!ix) return;
    free(ix->head); free(ix->start);
    for (int c = 0; c < 256; ++c) { free(ix->crun[c]); free(ix->ccum[c]); }
    free(ix->pred); free(ix->samples_last); free(ix->pred_to_run);
    free(ix->wstarts); free(ix->wends); free(ix->widxs); free(ix->arr);
    free(ix);
}

static u64* dup64(const u64* a, u64 m) {
    u64* p = (u64*) malloc((m ? m : 1) * sizeof(u64));
    if (m) memcpy(p, a, m * sizeof(u64));
    return p;
}

/* ---- rle_string ------------------------------------------------------- */

/* rle_string::rank(i,c), include/rle_string.hpp:131-161: number of c in BWT[0,i) */
u64 orc_rank(const orc_index* ix, u64 i, uint8_t c) {
    if (ix->nc[c] == 0) return 0;                         /* :134  letter does not exist */
    if (i == ix->n) return ix->ccum[c][ix->nc[c]];        /* :135  runs_per_letter[c].size() */
    /* :136-151 block lookup + scan of <=B runs == the run containing position i */
    u64 run = upper_bound(ix->start, ix->R, i) - 1;
    u64 dist = i - ix->start[run];
    u64 rk = lower_bound(ix->crun[c], ix->nc[c], run);    /* :155 run_heads.rank(run,c) */
    u64 tail = (ix->head[run] == c) ? dist : 0;           /* :157 */
    if (rk == 0) return tail;                             /* :160 */
    return ix->ccum[c][rk] + tail;                        /* :161 select(rk-1)+1 == length of first rk c-runs */
}

/* rle_string::operator[] / run_of_position, include/rle_string.hpp:99-102,166-186 */
u64 orc_run_of_position(const orc_index* ix, u64 i) {
    return upper_bound(ix->start, ix->R, i) - 1;
}
uint8_t orc_access(const orc_index* ix, u64 i) { return ix->head[orc_run_of_position(ix, i)]; }

/* rle_string::select(i,c), include/rle_string.hpp:107-126: position of the i-th (0-based) c */
u64 orc_select(const orc_index* ix, u64 i, uint8_t c) {
    u64 j = upper_bound(ix->ccum[c], ix->nc[c] + 1, i) - 1;   /* :111 runs_per_letter[c].rank(i): c-run holding it */
    u64 before = i - ix->ccum[c][j];                          /* :114 */
    u64 r = ix->crun[c][j];                                   /* :116 run_heads.select(j,c) */
    return ix->start[r] + before;                             /* :120-125 */
}

/* ---- RowBowt ------------------------------------------------------------ */

/* RowBowt::LF(range,c), include/rowbowt.hpp:74-88 */
static void LF(orc_index* ix, u64* lo, u64* hi, uint8_t c) {
    ix->lf_steps++;
    if ((c == 255 && ix->F[c] == ix->n) || ix->F[c] >= ix->F[c + 1]) { *lo = 1; *hi = 0; return; }
    u64 c_before = orc_rank(ix, *lo, c);
    u64 c_inside = orc_rank(ix, *hi + 1, c) - c_before;
    if (c_inside == 0) { *lo = 1; *hi = 0; return; }
    u64 l = ix->F[c] + c_before;
    *lo = l; *hi = l + c_inside - 1;
}

/* RowBowt::find_range, include/rowbowt.hpp:121-131 (no ftab: rb_align never loads one) */
void orc_find_range(orc_index* ix, const uint8_t* q, u64 m, u64* lo, u64* hi) {
    *lo = 0; *hi = ix->n - 1;                                  /* full_range :115-118 */
    for (u64 i = 0; i < m && *hi >= *lo; ++i) LF(ix, lo, hi, q[m - i - 1]);
}

/* RowBowt::LF_w_loc, include/rowbowt.hpp:555-573 */
static void LF_w_loc(orc_index* ix, u64* lo, u64* hi, uint8_t c, u64* k) {
    u64 olo = *lo, ohi = *hi;
    (void) olo;
    LF(ix, lo, hi, c);
    if (*lo <= *hi) {
        if (orc_access(ix, ohi) == c) {                        /* :559 trivial case */
            *k = *k - 1;
        } else {                                               /* :562-566 */
            u64 rnk = orc_rank(ix, ohi, c) - 1;
            u64 j = orc_select(ix, rnk, c);
            u64 run_of_j = orc_run_of_position(ix, j);
            *k = ix->samples_last[run_of_j];
        }
    } else { *lo = 1; *hi = 0; *k = 0; }                       /* :569-571 */
}

/* ToeholdSA::get_last_run_sample, include/toehold_sa.hpp:97-99 */
u64 orc_last_run_sample(const orc_index* ix) { return (ix->samples_last[ix->r - 1] + 1) % ix->n; }

/* RowBowt::find_range_w_toehold, include/rowbowt.hpp:169-184 */
void orc_find_range_w_toehold(orc_index* ix, const uint8_t* q, u64 m, u64* lo, u64* hi, u64* k) {
    if (!ix->has_tsa) { *lo = 1; *hi = 0; *k = 0; return; }   /* :171 default LFData; ssamp printed is unspecified */
    *lo = 0; *hi = ix->n - 1;
    *k = orc_last_run_sample(ix);
    for (u64 i = 0; i < m; ++i) {
        LF_w_loc(ix, lo, hi, q[m - i - 1], k);
        if (*hi < *lo) { *lo = 1; *hi = 0; *k = 0; return; }   /* :176-179 lf.clear() */
    }
}

/* ToeholdSA::phi, include/toehold_sa.hpp:56-72 */
u64 orc_phi(const orc_index* ix, u64 i) {
    u64 rk = lower_bound(ix->pred, ix->r, i);                  /* pred_.rank(i): ones strictly below i */
    u64 jr = rk == 0 ? ix->r - 1 : rk - 1;                     /* predecessor_rank_circular, sparse_sd_vector.hpp:141-143 */
    u64 j = ix->pred[jr];
    u64 delta = j < i ? i - j : i + 1;
    u64 prev_sample = ix->samples_last[ix->pred_to_run[jr] - 1];
    return (prev_sample + delta) % ix->n;
}

/* ToeholdSA::locate_range, include/toehold_sa.hpp:37-49; returns number written (<= cap) */
u64 orc_locate_range(const orc_index* ix, u64 l, u64 r, u64 k, u64 max_hits, u64* out, u64 cap) {
    u64 n_occ = r >= l ? (r - l) + 1 : 0;
    if (n_occ > max_hits) n_occ = max_hits;
    u64 k1 = k, w = 0;
    if (n_occ > 0) {
        if (w < cap) out[w] = k1;
        w++;
        for (u64 i = 1; i < n_occ; ++i) { k1 = orc_phi(ix, k1); if (w < cap) out[w] = k1; w++; }
    }
    return w;
}

/* ---- rle_window_arr ------------------------------------------------------- */
/* clamped rank/select helpers, pfbwt-f/include/rle_window_array.hpp:202-232 */
static u64 run_starts_rank(const orc_index* ix, u64 i) {       /* #starts <= i (clamped) */
    return i + 1 >= ix->sz_starts ? ix->nstarts : lower_bound(ix->wstarts, ix->nstarts, i + 1);
}
static u64 run_starts_select(const orc_index* ix, u64 i) {     /* 1-based */
    return i > ix->nstarts ? ix->sz_starts : ix->wstarts[i - 1];
}
static u64 run_ends_rank(const orc_index* ix, u64 i) {         /* #ends < i (clamped) */
    return i > ix->sz_ends - 1 ? lower_bound(ix->wends, ix->nends, ix->sz_ends - 1)
                               : lower_bound(ix->wends, ix->nends, i);
}
static u64 run_ends_select(const orc_index* ix, u64 i) {
    return i > ix->nends ? ix->sz_ends : ix->wends[i - 1];
}
static u64 arr_idxs_select(const orc_index* ix, u64 i) {
    return i > ix->nidxs ? ix->sz_idxs : ix->widxs[i - 1];
}
/* arr_at_, :236-243 (appends; no clearing) */
static u64 arr_at(const orc_index* ix, u64 i, u64* out, u64 cap, u64 w) {
    u64 s = arr_idxs_select(ix, i + 1), e = arr_idxs_select(ix, i + 2);
    for (u64 t = s; t < e; ++t) { if (w < cap) out[w] = ix->arr[t]; w++; }
    return w;
}
/* rle_window_arr::at, :114-120 */
u64 orc_markers_at(const orc_index* ix, u64 i, u64* out, u64 cap) {
    if (!ix->has_ma) return 0;
    u64 srank = run_starts_rank(ix, i), erank = run_ends_rank(ix, i);
    if (srank != erank + 1) return 0;
    return arr_at(ix, srank - 1, out, cap, 0);
}
/* rle_window_arr::at_range, :130-154; returns number of words (may exceed cap; only cap written) */
u64 orc_markers_at_range(const orc_index* ix, u64 s, u64 e, u64* out, u64 cap) {
    if (!ix->has_ma) return 0;
    u64 e_rs_rank = run_starts_rank(ix, e);
    if (e_rs_rank == 0) return 0;
    u64 e_rs_pos = run_starts_select(ix, e_rs_rank);
    if (e_rs_pos <= s) {
        u64 e_re_pos = run_ends_select(ix, e_rs_rank);
        return e_re_pos >= s ? arr_at(ix, e_rs_rank - 1, out, cap, 0) : 0;
    }
    u64 s_rs_rank = run_starts_rank(ix, s);
    s_rs_rank = s_rs_rank ? s_rs_rank : 1;
    u64 s_rs_pos = run_starts_select(ix, s_rs_rank);
    u64 s_re_rank = run_ends_rank(ix, s);
    s_re_rank = s_re_rank ? s_re_rank : 1;
    u64 s_re_pos = run_ends_select(ix, s_re_rank);
    u64 start_idx = s_rs_pos > s_re_pos ? s_rs_rank - 1 : s_rs_rank;
    u64 w = 0;
    for (u64 i = start_idx; i < e_rs_rank; ++i) w = arr_at(ix, i, out, cap, w);
    return w;
}

/* ---- construction ------------------------------------------------------------ */

orc_index* orc_create(u64 n, u64 R, const uint8_t* heads, const u64* lens) {
    orc_index* ix = (orc_index*) calloc(1, sizeof(orc_index));
    ix->n = n; ix->R = R;
    ix->head = (uint8_t*) malloc(R ? R : 1);
    memcpy(ix->head, heads, R);
    ix->start = (u64*) malloc((R + 1) * sizeof(u64));
    u64 p = 0;
    for (u64 j = 0; j < R; ++j) { ix->start[j] = p; p += lens[j]; ix->nc[heads[j]]++; }
    ix->start[R] = p;
    if (p != n) { orc_free(ix); return NULL; }
    u64 fill[256];
    for (int c = 0; c < 256; ++c) {
        ix->crun[c] = (u64*) malloc((ix->nc[c] ? ix->nc[c] : 1) * sizeof(u64));
        ix->ccum[c] = (u64*) calloc(ix->nc[c] + 1, sizeof(u64));
        fill[c] = 0;
    }
    for (u64 j = 0; j < R; ++j) {
        uint8_t c = heads[j];
        u64 k = fill[c]++;
        ix->crun[c][k] = j;
        ix->ccum[c][k + 1] = ix->ccum[c][k] + lens[j];
    }
    /* build_f, include/rowbowt.hpp:770-778: f_[i+1] = sum_{c<=i} rank(n,c), i in 0..254 */
    ix->F[0] = 0;
    for (int i = 0; i < 255; ++i) ix->F[i + 1] = ix->F[i] + orc_rank(ix, n, (uint8_t) i);
    ix->F[256] = n;    /* the reference reads f_[256] out of bounds for c==255; defined here */
    return ix;
}

void orc_set_tsa(orc_index* ix, u64 r, const u64* pred, const u64* samples_last, const u64* pred_to_run) {
    ix->r = r;
    ix->pred = dup64(pred, r);
    ix->samples_last = dup64(samples_last, r);
    ix->pred_to_run = dup64(pred_to_run, r);
    ix->has_tsa = 1;
}

void orc_set_markers(orc_index* ix, u64 sz_starts, u64 sz_ends, u64 sz_idxs, u64 nstarts, const u64* starts,
                     u64 nends, const u64* ends, u64 nidxs, const u64* idxs, u64 arr_size, const u64* arr) {
    ix->sz_starts = sz_starts; ix->sz_ends = sz_ends; ix->sz_idxs = sz_idxs;
    ix->nstarts = nstarts; ix->nends = nends; ix->nidxs = nidxs;
    ix->wstarts = dup64(starts, nstarts); ix->wends = dup64(ends, nends); ix->widxs = dup64(idxs, nidxs);
    ix->arr = dup64(arr, arr_size); ix->arr_size = arr_size;
    ix->has_ma = 1;
}

u64 orc_F(const orc_index* ix, int c) { return ix->F[c]; }
u64 orc_lf_steps(const orc_index* ix) { return ix->lf_steps; }

/* ---- batch drivers (what rb_report does per read, src/rb_align.cpp:118-145) ----- */

void orc_find_ranges(orc_index* ix, const uint8_t* bases, const u64* offs, u64 nreads, int toehold,
                     u64* lo, u64* hi, u64* k) {
    for (u64 i = 0; i < nreads; ++i) {
        const uint8_t* q = bases + offs[i];
        u64 m = offs[i + 1] - offs[i];
        if (toehold) orc_find_range_w_toehold(ix, q, m, &lo[i], &hi[i], &k[i]);
        else { orc_find_range(ix, q, m, &lo[i], &hi[i]); if (k) k[i] = 0; }
    }
}

/* ---- rb_markers: greedy-seeding marker genotyping ------------------- */

/* One fn(range, (q.first, q.second), mbuf) call of get_markers_greedy_seeding as the rb_markers worker
 * records it (MarkerSeed, src/rb_markers.cpp:255-275 and out_fn :356-373). */
typedef struct {
    u64 lo, hi;          /* p */
    u64 mk_off;          /* first word of this seed in the words buffer */
    uint32_t qstart;     /* MarkerSeed::query_start, low 32 bits (size_t(-1) -> 0xFFFFFFFF) */
    uint32_t qlen;       /* MarkerSeed::query_len */
    uint32_t mk_raw;     /* |mbuf| handed to fn (0 when range_size < min_range) */
    uint32_t mk_cnt;     /* after std::sort(marker_cmp) + std::unique */
} orc_seed;

/* RowBowt::search_ftab, include/rowbowt.hpp:745-758, over a dense 4^k table in build_ftab's enumeration
 * (base i of the k-mer -> bits 2i of the key); absent k-mers hold (1,0).  Returns 1 when found. */
static int ftab_find(u64 k, const u64* ft_lo, const u64* ft_hi, const uint8_t* s, u64* lo, u64* hi) {
    u64 key = 0;
    for (u64 i = 0; i < k; ++i) {
        int c = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : s[i] == 'T' ? 3 : -1;
        if (c < 0) return 0;
        key |= (u64) c << (2 * i);
    }
    if (ft_lo[key] > ft_hi[key]) return 0;
    *lo = ft_lo[key]; *hi = ft_hi[key];
    return 1;
}

typedef struct {
    orc_seed* seeds; u64 seed_cap, n_seeds;
    u64* words; u64 word_cap, n_words;
    u64 seed_first_word;     /* where the current seed's mbuf starts */
} greedy_out;

/* update_mbuf, include/rowbowt.hpp:437-441: mbuf = markers_at(r, mbuf) appends at_range(r) */
static void update_mbuf(const orc_index* ix, greedy_out* o, u64 lo, u64 hi, u64 max_range) {
    if (hi - lo + 1 <= max_range) {
        u64 room = o->n_words < o->word_cap ? o->word_cap - o->n_words : 0;
        u64 w = orc_markers_at_range(ix, lo, hi, room ? o->words + o->n_words : NULL, room);
        o->n_words += w;
    }
}

/* marker_cmp, src/rb_markers.cpp:243-251 (seq, pos, allele; pfbwt-f/include/marker.hpp) */
static int marker_less(u64 a, u64 b) {
    const u64 SEQ_MASK = 0x0FFFF00000000000ull, POS_MASK = 0x00000FFFFFFFFFFFull;
    u64 sa = (a & SEQ_MASK) >> 46, sb = (b & SEQ_MASK) >> 46, pa = a & POS_MASK, pb = b & POS_MASK;
    if (sa == sb && pa == pb) return (a >> 60) < (b >> 60);
    if (sa == sb) return pa < pb;
    return sa < sb;
}

/* out_fn of the worker, src/rb_markers.cpp:356-373, for strand `rev` of a read of `m` bases; then mbuf.clear() */
static void emit_seed(greedy_out* o, u64 lo, u64 hi, u64 qfirst, u64 qlast, int rev, u64 m, u64 min_range) {
    u64 first = o->seed_first_word, raw = o->n_words - first;
    if (!(hi < lo)) {                                            /* :365 empty ranges are not reported */
        u64 range_size = hi - lo + 1;
        u64 qstart = rev ? m - qfirst - 1 : qfirst;              /* :363 */
        u64 qlen = qlast - qfirst + 1;                           /* :364 */
        u64 cnt = 0;
        if (!(range_size >= min_range && raw)) raw = 0;          /* :366 */
        if (raw && o->n_words <= o->word_cap) {
            u64* w = o->words + first;
            for (u64 i = 1; i < raw; ++i) {                      /* std::sort(marker_cmp): insertion sort, ties by word */
                u64 x = w[i], j = i;
                while (j > 0 && (marker_less(x, w[j - 1]) || (!marker_less(w[j - 1], x) && x < w[j - 1]))) { w[j] = w[j - 1]; --j; }
                w[j] = x;
            }
            cnt = 1;
            for (u64 i = 1; i < raw; ++i) if (w[i] != w[cnt - 1]) w[cnt++] = w[i];      /* std::unique */
        }
        if (o->n_seeds < o->seed_cap) {
            orc_seed* s = &o->seeds[o->n_seeds];
            s->lo = lo; s->hi = hi; s->mk_off = first;
            s->qstart = (uint32_t) qstart; s->qlen = (uint32_t) qlen;
            s->mk_raw = (uint32_t) raw; s->mk_cnt = (uint32_t) cnt;
        }
        o->n_seeds++;
        o->n_words = first + raw;                                /* keep the raw slots: offsets stay a plain prefix sum */
    } else {
        o->n_words = first;
    }
    o->seed_first_word = o->n_words;
}

/* RowBowt::get_markers_greedy_seeding, include/rowbowt.hpp:406-482, one strand.  ft_k == 0: no ftab.
 * Returns -1 where the reference would throw/exit (ftab given and k - 1 > wsize :423-426, or m < k :431). */
static int greedy_seeding(orc_index* ix, const uint8_t* q, u64 m, u64 wsize, u64 max_range, u64 min_range,
                          u64 k, const u64* ft_lo, const u64* ft_hi, int rev, greedy_out* o) {
    if (k && k - 1 > wsize) return -1;
    if (k && m < k) return -1;
    const u64 flo = 0, fhi = ix->n - 1;                          /* full_range */
    u64 plo = flo, phi_ = fhi, lo = flo, hi = fhi, i = 0;
    if (k) {                                                     /* :430-433 */
        if (ftab_find(k, ft_lo, ft_hi, q + m - k, &lo, &hi)) i = k;
        plo = lo; phi_ = hi;
    }
    u64 window_ei = m, seed_ei = m;
    for (; i < m; ++i) {
        LF(ix, &lo, &hi, q[m - i - 1]);
        if (hi < lo) {                                           /* :444 the seed fails */
            if (seed_ei - (m - i) >= wsize) update_mbuf(ix, o, plo, phi_, max_range);
            emit_seed(o, plo, phi_, m - i, seed_ei - 1, rev, m, min_range);
            plo = flo; phi_ = fhi;
            seed_ei = m - i - 1; window_ei = m - i - 1;
            if (k && m - i - 1 >= k) {
                /* :454-464.  The loop there is written to "keep trying kmers, shifting left by one, until we find a
                 * match", but search_ftab answers a miss with (full_range(), 0) and the test that follows is
                 * range.first <= range.second, which the full range passes: the first iteration always breaks.  A
                 * missing k-mer therefore SKIPS its k bases and the seed goes on from the full range. */
                seed_ei = m - i - 1; window_ei = m - i - 1;
                lo = flo; hi = fhi;
                ftab_find(k, ft_lo, ft_hi, q + m - i - 1 - k, &lo, &hi);
                i += k; plo = lo; phi_ = hi;
            } else { lo = flo; hi = fhi; }
        } else {                                                 /* :468-474 window checkpoint */
            if (window_ei - (m - i - 1) >= wsize) { update_mbuf(ix, o, lo, hi, max_range); window_ei = m - i - 1; }
            plo = lo; phi_ = hi;
        }
    }
    if (hi >= lo && seed_ei - (m - i) >= wsize) update_mbuf(ix, o, lo, hi, max_range);      /* :478-480 */
    emit_seed(o, lo, hi, m - i, seed_ei - 1, rev, m, min_range);
    return 0;
}

/* seq_ntoa_table, src/rb_markers.cpp:135-152: acgtACGT -> ACGT, n/N -> A, everything else -> N */
static uint8_t ntoa(uint8_t c) {
    switch (c) {
        case 'A': case 'a': case 'N': case 'n': return 'A';
        case 'C': case 'c': return 'C';
        case 'G': case 'g': return 'G';
        case 'T': case 't': return 'T';
        default: return 'N';
    }
}
static uint8_t comp(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }   /* comp_tab on {A,C,G,T,N} */

/* The default worker of rb_markers (src/rb_markers.cpp:347-415) over a batch: for each read, forward strand then
 * reverse complement, seeds in generation order.  seed_off[2i+s] .. seed_off[2i+s+1] = seeds of read i strand s.
 * Returns 0, or -1 on the reference's error exits; *n_seeds / *n_words are the totals (may exceed the caps, in
 * which case only the caps were written: call again with larger buffers). */
int orc_rb_markers(orc_index* ix, const uint8_t* bases, const u64* offs, u64 nreads, u64 wsize, u64 max_range,
                   u64 min_range, u64 ft_k, const u64* ft_lo, const u64* ft_hi, u64* seed_off, orc_seed* seeds,
                   u64 seed_cap, u64* words, u64 word_cap, u64* n_seeds, u64* n_words) {
    greedy_out o = {seeds, seed_cap, 0, words, word_cap, 0, 0};
    uint8_t* buf = NULL;
    u64 cap = 0;
    int rc = 0;
    for (u64 r = 0; r < nreads && rc == 0; ++r) {
        u64 m = offs[r + 1] - offs[r];
        if (m > cap) { free(buf); cap = 2 * m + 16; buf = (uint8_t*) malloc(2 * cap); }
        uint8_t *fwd = buf, *rv = buf + cap;
        for (u64 j = 0; j < m; ++j) fwd[j] = ntoa(bases[offs[r] + j]);
        for (u64 j = 0; j < m; ++j) rv[j] = comp(fwd[m - 1 - j]);       /* KSeqString::revc_in_place :179-193 */
        seed_off[2 * r] = o.n_seeds;
        rc = greedy_seeding(ix, fwd, m, wsize, max_range, min_range, ft_k, ft_lo, ft_hi, 0, &o);
        seed_off[2 * r + 1] = o.n_seeds;
        if (rc == 0) rc = greedy_seeding(ix, rv, m, wsize, max_range, min_range, ft_k, ft_lo, ft_hi, 1, &o);
    }
    seed_off[2 * nreads] = o.n_seeds;
    free(buf);
    *n_seeds = o.n_seeds; *n_words = o.n_words;
    return rc;
}
