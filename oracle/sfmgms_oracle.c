/*
 * sfmgms_oracle.c — CPU restatement of the BF-Hamming + GMS matching path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * path (sfm_gms_b200/, include/sfmgms.h) never links, imports or falls back to it.
 *
 * What it restates (all reference paths are relative to /root/reference):
 *   stage 1  cv::BFMatcher(NORM_HAMMING, crossCheck=false)::match
 *            call site SfM-GMS/SfM-GMS/FeatureMatchUtil.cpp:66-68 (also :22-23, :40-41).
 *            The arithmetic lives in OpenCV features2d 4.5.x (third party, binary-only in the
 *            reference: SfM-GMS/lib/opencv_world45{1,2}.lib).  Semantics pinned against the
 *            executable cv2 4.13 BFMatcher in tests/test_oracle.py and tests/golden/.
 *   stage 2  cv::xfeatures2d::matchGMS  (OpenCV-contrib xfeatures2d 4.5.2, modules/xfeatures2d/src/gms.cpp)
 *            call sites FeatureMatchUtil.cpp:69 (true,true), DisparityUtil.cpp:149,299 (defaults).
 *            Third-party, binary-only in the reference (SfM-GMS/bin/opencv_xfeatures2d452.dll);
 *            the algorithm below follows the disassembly-verified specification in
 *            /root/repo/SURVEY.md Appendix A (virtual addresses quoted per function).
 *
 * PARITY PINNING: stage 1 is pinned by an executable reference (cv2).  Stage 2 is pinned by the
 * reference's OWN MACHINE CODE: oracle/dllref/gms_dll_host.c maps the vendored DLL
 * (SfM-GMS/bin/opencv_xfeatures2d452.dll) and calls its exported matchGMS (@VA 0x180048280), its
 * GMSMatcher ctor/setScale/run and its grid-index leaf functions; tests/golden/make_gms_dll_golden.py
 * records those outputs (tests/golden/gms_dll.npz) and tests/test_gms_dll.py checks this file
 * against them — masks, all 40 per-hypothesis masks, cell indices on edge-of-cell f32 values, and
 * the ROT / SCALE tables — plus, where the DLL is present, a live randomised comparison.  The older
 * pins stay as regressions: the survey's numpy anchors and the micro-cases.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every f32/f64 op is separately rounded,
 * exactly like the divss/mulss/addsd/divsd/sqrtsd/mulsd sequence in the DLL).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GRID_L 20 /* left grid 20x20: DLL @VA 0x180046ac6 */

/* ROT[8][9], 1-based, DLL .rdata 0x18012f520. */
static const int ROT[8][9] = {
    {1, 2, 3, 4, 5, 6, 7, 8, 9}, {4, 1, 2, 7, 5, 3, 8, 9, 6}, {7, 4, 1, 8, 5, 2, 9, 6, 3},
    {8, 7, 4, 9, 5, 1, 6, 3, 2}, {9, 8, 7, 6, 5, 4, 3, 2, 1}, {6, 9, 8, 3, 5, 7, 2, 1, 4},
    {3, 6, 9, 2, 5, 8, 1, 4, 7}, {2, 3, 6, 1, 5, 9, 4, 7, 8}};

/* cvFloor / cvRound as OpenCV defines them (SURVEY Appendix A "Helpers"). */
static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
static inline int cv_floor_d(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_round_d(double v) { return (int)lrint(v); } /* cvtsd2si: half-to-even */

/* ------------------------------------------------------------------------------------------
 * Stage 1: brute-force Hamming NN.  SURVEY §8(a1) / Appendix B.
 * strict '<' scan in increasing train index => lowest trainIdx on ties.
 * Returns 0, or -1 on bad arguments (nt >= 2^18 mirrors OpenCV's IMGIDX_ONE assert).
 * nt == 0 => no matches are produced (caller sees n_matches = 0); outputs untouched.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *q, *t;
    int nt, desc_bytes, i0, i1;
    int32_t *train_idx, *dist;
} bf_job_t;

static void* bf_worker(void* arg) {
    const bf_job_t* jb = (const bf_job_t*)arg;
    const int desc_bytes = jb->desc_bytes;
    const int words = desc_bytes / 8, tail = desc_bytes % 8;
    for (int i = jb->i0; i < jb->i1; ++i) {
        const uint8_t* a = jb->q + (size_t)i * desc_bytes;
        int best = 0x7fffffff, bestj = -1;
        for (int j = 0; j < jb->nt; ++j) {
            const uint8_t* b = jb->t + (size_t)j * desc_bytes;
            int d = 0;
            for (int w = 0; w < words; ++w) {
                uint64_t x, y;
                memcpy(&x, a + 8 * w, 8);
                memcpy(&y, b + 8 * w, 8);
                d += __builtin_popcountll(x ^ y);
            }
            for (int k = 0; k < tail; ++k)
                d += __builtin_popcount((unsigned)(a[8 * words + k] ^ b[8 * words + k]));
            if (d < best) { best = d; bestj = j; } /* strict '<': lowest trainIdx on ties */
        }
        jb->train_idx[i] = bestj;
        jb->dist[i] = best;
    }
    return NULL;
}

static int g_threads = 1;
/* Threads used by oracle_bf_hamming (OpenCV's batchDistance is parallel_for_ over query rows;
 * GMS is single-threaded in OpenCV and stays so here). */
void oracle_set_num_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int oracle_num_threads(void) { return g_threads; }

int oracle_bf_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int desc_bytes,
                      int32_t* train_idx, int32_t* dist, int* n_matches) {
    if (nq < 0 || nt < 0 || desc_bytes <= 0 || nt >= (1 << 18)) return -1;
    if (nt == 0) { if (n_matches) *n_matches = 0; return 0; }
    int nth = g_threads;
    if (nth > nq) nth = nq > 0 ? nq : 1;
    bf_job_t jobs[256];
    pthread_t th[256];
    for (int k = 0; k < nth; ++k) {
        bf_job_t jb = {q, t, nt, desc_bytes, (int)((long)nq * k / nth),
                       (int)((long)nq * (k + 1) / nth), train_idx, dist};
        jobs[k] = jb;
    }
    for (int k = 1; k < nth; ++k) pthread_create(&th[k], NULL, bf_worker, &jobs[k]);
    bf_worker(&jobs[0]);
    for (int k = 1; k < nth; ++k) pthread_join(th[k], NULL);
    if (n_matches) *n_matches = nq;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Stage 2: GMS.  One struct mirrors the GMSMatcher object (DLL ctor @VA 0x180046900).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int n;                 /* mNumberMatches */
    const float* p1;       /* normalised points image 1 (x,y) */
    const float* p2;       /* normalised points image 2 */
    const int32_t* mq;     /* match.first  (queryIdx) */
    const int32_t* mt;     /* match.second (trainIdx) */
    int wr, hr, gr;        /* right grid (setScale) */
    int32_t* hist;         /* GL x GR motion statistics */
    int32_t* cnt;          /* GL points per left cell */
    int32_t* cp;           /* GL cell pairs */
    int32_t* pl;           /* per match: pair.first  */
    int32_t* pr;           /* per match: pair.second */
    uint8_t* mask;         /* per match inlier mask of the current run */
    int32_t nbl[GRID_L * GRID_L][9];
    int32_t* nbr;          /* GR x 9 */
    double factor;
} gms_t;

/* neighbors(W,H): DLL @VA 0x180048030. */
static void build_nb9(int32_t* nb, int w, int h) {
    for (int idx = 0; idx < w * h; ++idx) {
        int x = idx % w, y = idx / w;
        for (int k = 0; k < 9; ++k) nb[idx * 9 + k] = -1;
        for (int yi = -1; yi <= 1; ++yi)
            for (int xi = -1; xi <= 1; ++xi) {
                int xx = x + xi, yy = y + yi;
                if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
                nb[idx * 9 + (yi + 1) * 3 + (xi + 1)] = xx + yy * w;
            }
    }
}

/* SCALE table, DLL .data 0x1802c5008; [2],[3] are sqrt() results at static init. */
static double scale_ratio(int s) {
    switch (s) {
        case 0: return 1.0;
        case 1: return 0.5;
        case 2: return 1.0 / sqrt(2.0);
        case 3: return sqrt(2.0);
        default: return 2.0;
    }
}

/* setScale: DLL @VA 0x180048c10.  => (wr,gr) = (20,400)(10,100)(14,196)(28,784)(40,1600). */
static void set_scale(gms_t* g, int s) {
    g->wr = cv_round_d((double)GRID_L * scale_ratio(s));
    g->hr = cv_round_d((double)GRID_L * scale_ratio(s));
    g->gr = g->wr * g->hr;
    build_nb9(g->nbr, g->wr, g->hr);
}

/* getGridIndexLeft: DLL @VA 0x180047bc0.  f32 multiply, f64 add of 0.5 on the shifted axis. */
static int left_idx(const float* p, int type) {
    float fx = (float)GRID_L * p[0];
    float fy = (float)GRID_L * p[1];
    int x = (type == 2 || type == 4) ? cv_floor_d((double)fx + 0.5) : cv_floor_f(fx);
    int y = (type == 3 || type == 4) ? cv_floor_d((double)fy + 0.5) : cv_floor_f(fy);
    if (x >= GRID_L || y >= GRID_L) return -1;
    return x + y * GRID_L;
}

/* getGridIndexRight: DLL @VA 0x180047d60 (no bounds check in the reference; the supported
 * domain 0<=x<w, 0<=y<h keeps it inside [0,gr), validated by oracle_gms before any run). */
static int right_idx(const gms_t* g, const float* p) {
    return cv_floor_f((float)g->wr * p[0]) + cv_floor_f((float)g->hr * p[1]) * g->wr;
}

/* verifyCellPairs: DLL @VA 0x180048d10. */
static void verify(gms_t* g, int rot) {
    const int* rp = ROT[rot - 1];
    const int GL = GRID_L * GRID_L;
    for (int i = 0; i < GL; ++i) {
        const int32_t* row = g->hist + (size_t)i * g->gr;
        long sum = 0;
        for (int j = 0; j < g->gr; ++j) sum += row[j]; /* cv::sum(row) @VA 0x180048da0 */
        if (sum == 0) { g->cp[i] = -1; continue; }
        int maxv = 0;
        for (int j = 0; j < g->gr; ++j)
            if (row[j] > maxv) { g->cp[i] = j; maxv = row[j]; }
        int score = 0, num = 0;
        double thresh = 0.0;
        for (int k = 0; k < 9; ++k) {
            int ll = g->nbl[i][k];
            int rr = g->nbr[g->cp[i] * 9 + rp[k] - 1];
            if (ll == -1 || rr == -1) continue;
            score += g->hist[(size_t)ll * g->gr + rr];
            thresh += (double)g->cnt[ll];
            num++;
        }
        thresh = g->factor * sqrt(thresh / (double)num); /* divsd, sqrtsd, mulsd */
        if ((double)score < thresh) g->cp[i] = -2;
    }
}

/* run: DLL @VA 0x180048630 (assignMatchPairs inlined @VA 0x1800489e0). Returns popcount(mask). */
static int run(gms_t* g, int rot) {
    const int GL = GRID_L * GRID_L;
    memset(g->mask, 0, (size_t)g->n);
    for (int i = 0; i < g->n; ++i) { g->pl[i] = 0; g->pr[i] = 0; }
    for (int type = 1; type <= 4; ++type) {
        memset(g->hist, 0, sizeof(int32_t) * (size_t)GL * g->gr);
        for (int i = 0; i < GL; ++i) { g->cp[i] = -1; g->cnt[i] = 0; }
        for (int i = 0; i < g->n; ++i) {
            int l = g->pl[i] = left_idx(g->p1 + 2 * (size_t)g->mq[i], type);
            int r;
            if (type == 1) r = g->pr[i] = right_idx(g, g->p2 + 2 * (size_t)g->mt[i]);
            else r = g->pr[i];
            if (l < 0 || r < 0) continue;
            g->hist[(size_t)l * g->gr + r]++;
            g->cnt[l]++;
        }
        verify(g, rot);
        for (int i = 0; i < g->n; ++i)
            if (g->pl[i] >= 0 && g->cp[g->pl[i]] == g->pr[i]) g->mask[i] = 1;
    }
    int c = 0;
    for (int i = 0; i < g->n; ++i) c += g->mask[i];
    return c;
}

/*
 * matchGMS / GMSMatcher::getInlierMask (DLL @VA 0x180048280 / 0x180047dc0).
 *
 * kp*_xy: keypoint pixel coordinates, interleaved x,y (stride in floats given by kp_stride,
 *         2 for packed; 7 for an array of cv::KeyPoint).
 * mask:   n_matches bytes (0/1).  *mask_len is n_matches, or 0 when rotation/scale search was
 *         requested and every hypothesis scored 0 (the reference leaves the vector untouched).
 * hyp_counts (optional, 40 ints): inlier count per hypothesis in scale-major order, -1 = not run.
 * best_hyp (optional): scale*8 + (rot-1) of the winning hypothesis (-1 if none).
 * Returns 0; -2 if a referenced keypoint is outside [0,w)x[0,h) (UB in the reference);
 * -3 if a match index is out of range (UB in the reference); -1 on bad arguments.
 */
int oracle_gms_ex(int w1, int h1, int w2, int h2, const float* kp1_xy, int n1, int kp1_stride,
               const float* kp2_xy, int n2, int kp2_stride, const int32_t* query_idx,
               const int32_t* train_idx, int idx_stride, int n_matches, int with_rotation,
               int with_scale, double threshold_factor, uint8_t* mask, int* mask_len,
               int* n_inliers, int* hyp_counts, int* best_hyp, uint8_t* all_masks);

int oracle_gms(int w1, int h1, int w2, int h2, const float* kp1_xy, int n1, int kp1_stride,
               const float* kp2_xy, int n2, int kp2_stride, const int32_t* query_idx,
               const int32_t* train_idx, int idx_stride, int n_matches, int with_rotation,
               int with_scale, double threshold_factor, uint8_t* mask, int* mask_len,
               int* n_inliers, int* hyp_counts, int* best_hyp) {
    return oracle_gms_ex(w1, h1, w2, h2, kp1_xy, n1, kp1_stride, kp2_xy, n2, kp2_stride, query_idx, train_idx,
                         idx_stride, n_matches, with_rotation, with_scale, threshold_factor, mask, mask_len,
                         n_inliers, hyp_counts, best_hyp, NULL);
}

/* Tables and leaf functions, exported so that tests can hold them against the DLL's own. */
void oracle_gms_tables(int32_t rot[72], double scale[5]) {
    for (int r = 0; r < 8; ++r) for (int k = 0; k < 9; ++k) rot[r * 9 + k] = ROT[r][k];
    for (int s = 0; s < 5; ++s) scale[s] = scale_ratio(s);
}
void oracle_gms_grid_left(const float* norm_xy, long n, int type, int32_t* out) {
    for (long i = 0; i < n; ++i) out[i] = left_idx(norm_xy + 2 * i, type);
}
void oracle_gms_grid_right(const float* norm_xy, long n, int wr, int hr, int32_t* out) {
    gms_t g; g.wr = wr; g.hr = hr;
    for (long i = 0; i < n; ++i) out[i] = right_idx(&g, norm_xy + 2 * i);
}
int oracle_gms_right_grid(int s) { return cv_round_d((double)GRID_L * scale_ratio(s)); }

/* all_masks (optional): [40][n_matches] bytes, the mask of EVERY hypothesis that was run (scale-major). */
int oracle_gms_ex(int w1, int h1, int w2, int h2, const float* kp1_xy, int n1, int kp1_stride,
               const float* kp2_xy, int n2, int kp2_stride, const int32_t* query_idx,
               const int32_t* train_idx, int idx_stride, int n_matches, int with_rotation,
               int with_scale, double threshold_factor, uint8_t* mask, int* mask_len,
               int* n_inliers, int* hyp_counts, int* best_hyp, uint8_t* all_masks) {
    if (n1 < 0 || n2 < 0 || n_matches < 0 || w1 <= 0 || h1 <= 0 || w2 <= 0 || h2 <= 0) return -1;
    const int GL = GRID_L * GRID_L;
    int rc = 0;
    gms_t g;
    memset(&g, 0, sizeof g);
    float* p1 = (float*)malloc(sizeof(float) * 2 * (size_t)(n1 ? n1 : 1));
    float* p2 = (float*)malloc(sizeof(float) * 2 * (size_t)(n2 ? n2 : 1));
    int32_t* mq = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_matches ? n_matches : 1));
    int32_t* mt = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_matches ? n_matches : 1));
    /* normalizePoints: DLL @VA 0x180048420 — f32 divss by (float)size. */
    for (int i = 0; i < n1; ++i) {
        p1[2 * i] = kp1_xy[(size_t)i * kp1_stride] / (float)w1;
        p1[2 * i + 1] = kp1_xy[(size_t)i * kp1_stride + 1] / (float)h1;
    }
    for (int i = 0; i < n2; ++i) {
        p2[2 * i] = kp2_xy[(size_t)i * kp2_stride] / (float)w2;
        p2[2 * i + 1] = kp2_xy[(size_t)i * kp2_stride + 1] / (float)h2;
    }
    for (int i = 0; i < n_matches; ++i) {
        mq[i] = query_idx[(size_t)i * idx_stride];
        mt[i] = train_idx[(size_t)i * idx_stride];
        if (mq[i] < 0 || mq[i] >= n1 || mt[i] < 0 || mt[i] >= n2) { rc = -3; goto done; }
        const float* a = kp1_xy + (size_t)mq[i] * kp1_stride;
        const float* b = kp2_xy + (size_t)mt[i] * kp2_stride;
        if (!(a[0] >= 0.f && a[0] < (float)w1 && a[1] >= 0.f && a[1] < (float)h1) ||
            !(b[0] >= 0.f && b[0] < (float)w2 && b[1] >= 0.f && b[1] < (float)h2)) {
            rc = -2;
            goto done;
        }
    }
    g.n = n_matches; g.p1 = p1; g.p2 = p2; g.mq = mq; g.mt = mt; g.factor = threshold_factor;
    g.hist = (int32_t*)malloc(sizeof(int32_t) * (size_t)GL * 1600);
    g.cnt = (int32_t*)malloc(sizeof(int32_t) * GL);
    g.cp = (int32_t*)malloc(sizeof(int32_t) * GL);
    g.pl = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_matches ? n_matches : 1));
    g.pr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_matches ? n_matches : 1));
    g.mask = (uint8_t*)malloc((size_t)(n_matches ? n_matches : 1));
    g.nbr = (int32_t*)malloc(sizeof(int32_t) * 1600 * 9);
    build_nb9(&g.nbl[0][0], GRID_L, GRID_L);

    if (hyp_counts) for (int k = 0; k < 40; ++k) hyp_counts[k] = -1;
    int best = 0, bh = -1, len = 0;
    if (!with_rotation && !with_scale) {
        set_scale(&g, 0);
        best = run(&g, 1);
        if (n_matches) memcpy(mask, g.mask, (size_t)n_matches);
        len = n_matches;
        bh = 0;
        if (hyp_counts) hyp_counts[0] = best;
        if (all_masks && n_matches) memcpy(all_masks, g.mask, (size_t)n_matches);
    } else {
        for (int s = 0; s < (with_scale ? 5 : 1); ++s) {
            set_scale(&g, s);
            for (int r = 1; r <= (with_rotation ? 8 : 1); ++r) {
                int c = run(&g, r);
                if (hyp_counts) hyp_counts[s * 8 + r - 1] = c;
                if (all_masks && n_matches) memcpy(all_masks + (size_t)(s * 8 + r - 1) * n_matches, g.mask, (size_t)n_matches);
                if (c > best) { /* strict >, first wins */
                    memcpy(mask, g.mask, (size_t)n_matches);
                    best = c; len = n_matches; bh = s * 8 + r - 1;
                }
            }
        }
    }
    if (mask_len) *mask_len = len;
    if (n_inliers) *n_inliers = best;
    if (best_hyp) *best_hyp = bh;
    free(g.hist); free(g.cnt); free(g.cp); free(g.pl); free(g.pr); free(g.mask); free(g.nbr);
done:
    free(p1); free(p2); free(mq); free(mt);
    return rc;
}

/* OpenCV core hal::normL2Sqr_(const float* a, const float* b, int n) (modules/core/src/norm.cpp), universal-
 * intrinsics build with 4 float lanes (the SSE baseline of the stock x86-64 packages): four vector accumulators,
 * 16 elements per iteration, d_k = t*t + d_k with mul and add rounded separately; lanes combined as
 * ((d0+d1)+d2)+d3, then v_reduce_sum = (s0+s2)+(s1+s3); the n%16 tail is added element by element.
 * Pinned: bit-exact against cv2 4.13 here (tests/test_oracle.py::test_l2_oracle_general_float_vs_cv2). */
static float norm_l2_sqr_cv(const float* a, const float* b, int n) {
    float acc[16];
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    int j = 0;
    for (; j <= n - 16; j += 16)
        for (int c = 0; c < 16; ++c) { float t = a[j + c] - b[j + c]; float m = t * t; acc[c] = m + acc[c]; }
    float s[4];
    for (int l = 0; l < 4; ++l) { float x = acc[l] + acc[4 + l]; x = x + acc[8 + l]; s[l] = x + acc[12 + l]; }
    float x = s[0] + s[2], y = s[1] + s[3];
    float d = x + y;
    for (; j < n; ++j) { float t = a[j] - b[j]; float m = t * t; d = d + m; }
    return d;
}

/* (f3) cv::BFMatcher(NORM_L2, crossCheck=false)::match on float descriptors (what the reference literally runs:
 * SIFT + BFMatcher::create(), FeatureMatchUtil.cpp:10, 66-68).  OpenCV: dist = sqrt(normL2Sqr_(a, b)) in float,
 * strict '<' scan on the float distance => lowest trainIdx among equal FLOAT distances. */
int oracle_bf_l2(const float* q, int nq, const float* t, int nt, int dim, int32_t* train_idx, float* dist, int* n_matches) {
    if (nq < 0 || nt < 0 || dim <= 0 || nt >= (1 << 18)) return -1;
    if (nt == 0) { if (n_matches) *n_matches = 0; return 0; }
    for (int i = 0; i < nq; ++i) {
        const float* a = q + (size_t)i * dim;
        float best = INFINITY;
        int bestj = -1;
        for (int j = 0; j < nt; ++j) {
            float d = sqrtf(norm_l2_sqr_cv(a, t + (size_t)j * dim, dim));
            if (d < best) { best = d; bestj = j; }
        }
        train_idx[i] = bestj;
        dist[i] = best;
    }
    if (n_matches) *n_matches = nq;
    return 0;
}

/* (f2) cross-check helper for the "next" row: mutual nearest neighbours, as
 * BFMatcher(NORM_HAMMING, crossCheck=true) does (FeatureMatchUtil.cpp:22 uses crossCheck=true
 * with NORM_L2): keep (i, j*) iff i is also the NN of j* when the roles are swapped. */
int oracle_bf_hamming_crosscheck(const uint8_t* q, int nq, const uint8_t* t, int nt,
                                 int desc_bytes, int32_t* train_idx, int32_t* dist,
                                 uint8_t* keep) {
    if (nq >= (1 << 18)) return -1;
    int n = 0;
    int rc = oracle_bf_hamming(q, nq, t, nt, desc_bytes, train_idx, dist, &n);
    if (rc) return rc;
    if (nt == 0 || nq == 0) return 0;
    int32_t* back = (int32_t*)malloc(sizeof(int32_t) * (size_t)nt);
    int32_t* bd = (int32_t*)malloc(sizeof(int32_t) * (size_t)nt);
    rc = oracle_bf_hamming(t, nt, q, nq, desc_bytes, back, bd, &n);
    for (int i = 0; i < nq; ++i) keep[i] = (back[train_idx[i]] == i);
    free(back); free(bd);
    return rc;
}

/* (f4) The 7x7 sigma-2 Gaussian ORB applies to every pyramid level before sampling the BRIEF tests (orb.cpp
 * detectAndCompute: GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on a SUB-matrix of the
 * pyramid, which skips OpenCV's 8-bit fixed-point Gaussian and runs the float separable filter).  Pinned against cv2
 * 4.13 (x86-64 AVX2 dispatch of imgproc filter.simd.hpp: RowVec_32f / SymmColumnVec_32f): float taps
 * k = (float)getGaussianKernel(7, 2); row pass sequential, s = x0*k0 then s = fma(x_i, k_i, s) in the first
 * 32*floor(w/32) columns (the vector loop) and s = s + x_i*k_i with two roundings in the remaining columns (its scalar
 * remainder); column pass symmetric, s = r3*k3, s = fma(r[3-d] + r[3+d], k[3-d], s) for d = 1..3 (two roundings
 * instead of the fma in the last w mod 4 columns, the column filter's scalar tail -- thin evidence: two events of the
 * randomised stress with edgeThreshold 0 and the float-source filter's visible tail; only reachable with
 * edgeThreshold < 13); result rounded half to even.  Evidence: descriptor bits of cv2.ORB on 7 full-size images x 20,000 keypoints and a
 * randomised live-cv2 stress (scripts/orb_stress.py: 0 mismatches in 14,714 trials); the float-source variant of the
 * same filter is visible directly through cv2.GaussianBlur(float32).  The loop width was settled on 319 images whose
 * blurred pyramids differ between a 32- and a 64-column vector loop: 32 columns 0 differing descriptor bits, 64
 * columns 80; fusing everywhere: 1-5 bits on about every second such image. */
static int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

int oracle_orb_blur7(const uint8_t* src, int w, int h, uint8_t* dst) {
    static const float k[7] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95c0p-3f, 0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};
    if (w <= 0 || h <= 0) return -1;
    float* rows = (float*)malloc(sizeof(float) * (size_t)w * (size_t)h);
    if (!rows) return -1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* r = src + (size_t)y * w;
            float s = (float)r[reflect101(x - 3, w)] * k[0];
            if (x < (w & ~31)) {      /* the 32-pixel vector loop of the uchar->float row filter fuses ... */
                for (int i = 1; i < 7; ++i) s = fmaf((float)r[reflect101(x - 3 + i, w)], k[i], s);
            } else {                  /* ... its scalar remainder rounds product and sum separately */
                for (int i = 1; i < 7; ++i) { float m = (float)r[reflect101(x - 3 + i, w)] * k[i]; s = s + m; }
            }
            rows[(size_t)y * w + x] = s;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float s = rows[(size_t)y * w + x] * k[3];
            for (int d = 1; d <= 3; ++d) {
                float pair = rows[(size_t)reflect101(y - d, h) * w + x] + rows[(size_t)reflect101(y + d, h) * w + x];
                if (x < (w & ~3)) s = fmaf(pair, k[3 - d], s);          /* vector loops of the column filter (8, 4 lanes) */
                else { float m = pair * k[3 - d]; s = s + m; }          /* its scalar tail: two roundings */
            }
            float r = nearbyintf(s);
            dst[(size_t)y * w + x] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
        }
    free(rows);
    return 0;
}
