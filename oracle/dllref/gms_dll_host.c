/* TEST INFRASTRUCTURE — never linked into, loaded by or called from the product (sfm_gms_b200/).
 *
 * Host for the reference's OWN machine code of cv::xfeatures2d::matchGMS.
 *
 * The only implementation of stage 2 (GMS) inside /root/reference is the vendored Windows binary
 *   /root/reference/SfM-GMS/bin/opencv_xfeatures2d452.dll   (OpenCV contrib 4.5.2, x86-64 PE32+)
 * that the reference links (`SfM-GMS/SfM-GMS/SfM-GMS.vcxproj:131,150`) and calls at
 * `SfM-GMS/SfM-GMS/FeatureMatchUtil.cpp:69`, `DisparityUtil.cpp:149,299`.  This file maps that DLL's
 * sections into memory on Linux (same ISA), applies its base relocations, points its import table at the
 * small stand-ins below (C runtime + the handful of cv::Mat members GMS uses: ctor/dtor, Mat::zeros, row
 * ROI, setTo(0), cv::sum) and then calls the DLL's functions through the Microsoft x64 calling convention:
 *   export  cv::xfeatures2d::matchGMS            @VA 0x180048280   (whole stage, as FeatureMatchUtil.cpp:69 does)
 *   GMSMatcher ctor 0x180046900 / dtor 0x180046d20 / setScale 0x180048c10 / run 0x180048630
 *   getGridIndexLeft 0x180047bc0 / getGridIndexRight 0x180047d60 (leaf functions)
 *   static initialiser of the scale table 0x1800010b0; tables .rdata 0x18012f520 (ROT), .data 0x1802c5008 (SCALE)
 * (addresses: SURVEY.md Appendix A).  Every arithmetic instruction that decides a GMS outcome — the f32
 * divisions and multiplications, the f64 +0.5, the floor, the histogram, the argmax, the f64
 * div/sqrt/mul threshold, the best-hypothesis loop — executes from the DLL's .text; the stand-ins only
 * allocate, zero, and add up one row of int32 (cv::sum, used by the DLL for the "row empty?" test only).
 *
 * Built by oracle/Makefile into oracle/_ref/libgms_dll_host.so (git-ignored).  No byte of the DLL is copied
 * into the repository: the path of the DLL is an argument of gmsdll_load() and it is read at run time, in the
 * build container only (the GPU box has no /root/reference; tests use the committed outputs
 * tests/golden/gms_dll_*.npz written by tests/golden/make_gms_dll_golden.py).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>
#include <sys/mman.h>

#define MS __attribute__((ms_abi))
#define IMAGE_BASE 0x180000000ULL

static uint8_t* g_img = NULL;
static uint64_t g_size = 0;
static char g_err[256];

static inline void* va(uint64_t a) { return g_img + (a - IMAGE_BASE); }

/* ------------------------------------------------------------------ cv::Mat stand-in (OpenCV 4.5.2 layout, 96 bytes) */
typedef struct UData { void* prevA; void* currA; int urefcount; int refcount; uint8_t* data; uint8_t* origdata; size_t size;
                       int flags; void* handle; void* userdata; int allocFlags; int mapcount; void* orig; } UData;
typedef struct Mat {
  int flags, dims, rows, cols;
  uint8_t* data; const uint8_t *datastart, *dataend, *datalimit;
  void* allocator; UData* u;
  int* size_p;          /* MatSize  */
  size_t* step_p;       /* MatStep  */
  size_t step_buf[2];
} Mat;
typedef struct Range { int start, end; } Range;
typedef struct InputArray { int flags; void* obj; int sz_w, sz_h; } InputArray;
typedef struct MatExpr { const void* op; int flags; int pad; Mat a, b, c; double alpha, beta; double s[4]; } MatExpr;

#define CV_32S 4
#define MAT_MAGIC 0x42FF0000
#define MAT_CONT 0x4000

static void mat_init(Mat* m) {
  memset(m, 0, sizeof *m);
  m->flags = MAT_MAGIC;
  m->size_p = &m->rows;
  m->step_p = m->step_buf;
}
static void mat_release(Mat* m) {
  if (m->u) {
    if (--m->u->refcount == 0) { free(m->u->origdata); free(m->u); }
  }
  m->u = NULL; m->data = NULL; m->datastart = m->dataend = m->datalimit = NULL;
  m->rows = m->cols = 0;
}
static void mat_create_zero(Mat* m, int rows, int cols, int type) {
  if (type != CV_32S) { fprintf(stderr, "gms_dll_host: Mat type %d not supported\n", type); abort(); }
  mat_release(m);
  size_t bytes = (size_t)rows * cols * 4;
  UData* u = (UData*)calloc(1, sizeof(UData));
  u->refcount = 1;
  u->origdata = u->data = (uint8_t*)calloc(bytes ? bytes : 1, 1);
  u->size = bytes;
  m->flags = MAT_MAGIC | MAT_CONT | type;
  m->dims = 2; m->rows = rows; m->cols = cols;
  m->data = u->data; m->datastart = u->data; m->dataend = m->datalimit = u->data + bytes;
  m->u = u;
  m->size_p = &m->rows; m->step_p = m->step_buf;
  m->step_buf[0] = (size_t)cols * 4; m->step_buf[1] = 4;
}

static MS Mat* s_mat_ctor(Mat* self) { mat_init(self); return self; }                      /* ??0Mat@cv@@QEAA@XZ */
static MS void s_mat_dtor(Mat* self) { mat_release(self); }                                /* ??1Mat@cv@@QEAA@XZ */
static MS Mat* s_mat_roi(Mat* self, const Mat* m, const Range* rr, const Range* cr) {      /* Mat(const Mat&, Range, Range) */
  mat_init(self);
  int r0 = 0, r1 = m->rows, c0 = 0, c1 = m->cols;
  if (!(rr->start == INT_MIN && rr->end == INT_MAX)) { r0 = rr->start; r1 = rr->end; }
  if (!(cr->start == INT_MIN && cr->end == INT_MAX)) { c0 = cr->start; c1 = cr->end; }
  if (r0 < 0 || r1 > m->rows || r0 > r1 || c0 < 0 || c1 > m->cols || c0 > c1) {
    fprintf(stderr, "gms_dll_host: ROI out of range\n"); abort();
  }
  self->flags = m->flags; self->dims = 2; self->rows = r1 - r0; self->cols = c1 - c0;
  self->step_buf[0] = m->step_p[0]; self->step_buf[1] = m->step_p[1];
  self->data = m->data + (size_t)r0 * m->step_p[0] + (size_t)c0 * 4;
  self->datastart = m->datastart; self->dataend = m->dataend; self->datalimit = m->datalimit;
  self->u = m->u; if (self->u) self->u->refcount++;
  if (self->cols != m->cols && self->rows > 1) self->flags &= ~MAT_CONT;
  return self;
}
/* MatExpr returned by Mat::zeros: the DLL then calls expr.op->assign(expr, dst, -1) through vtable slot 2 */
static MS void s_op_assign(const void* op, const MatExpr* e, Mat* dst, int type) {
  (void)op; (void)type;
  mat_create_zero(dst, e->a.rows, e->a.cols, e->flags);
}
static MS void s_op_trap(void) { fprintf(stderr, "gms_dll_host: unexpected MatOp virtual call\n"); abort(); }
static void* g_op_vtbl[16];
static struct { void** vptr; } g_op = { g_op_vtbl };
static MS MatExpr* s_mat_zeros(MatExpr* ret, int rows, int cols, int type) {              /* Mat::zeros(int,int,int) */
  memset(ret, 0, sizeof *ret);
  ret->op = &g_op; ret->flags = type;
  mat_init(&ret->a); mat_init(&ret->b); mat_init(&ret->c);
  ret->a.rows = rows; ret->a.cols = cols;   /* carried to s_op_assign; a owns nothing */
  return ret;
}
static const Mat* ia_mat(const InputArray* a) {
  int kind = a->flags & (31 << 16);
  if (kind != (1 << 16)) { fprintf(stderr, "gms_dll_host: InputArray kind 0x%x not a Mat\n", a->flags); abort(); }
  return (const Mat*)a->obj;
}
static MS Mat* s_mat_setTo(Mat* self, const InputArray* v, const InputArray* mask) {       /* Mat::setTo(value, mask) */
  (void)mask;
  /* GMS only ever calls setTo(0): check the scalar really is zero */
  const double* d = (const double*)v->obj;
  int kind = v->flags & (31 << 16);
  if (kind == (1 << 16)) { fprintf(stderr, "gms_dll_host: setTo(Mat) unsupported\n"); abort(); }
  if (d[0] != 0.0) { fprintf(stderr, "gms_dll_host: setTo(%g) unsupported\n", d[0]); abort(); }
  for (int r = 0; r < self->rows; r++) memset(self->data + (size_t)r * self->step_p[0], 0, (size_t)self->cols * 4);
  return self;
}
static MS double* s_cv_sum(double* ret, const InputArray* a) {                              /* cv::sum -> Scalar */
  const Mat* m = ia_mat(a);
  if ((m->flags & 0xFFF) != CV_32S) { fprintf(stderr, "gms_dll_host: sum of type %d\n", m->flags & 0xFFF); abort(); }
  double s = 0;
  for (int r = 0; r < m->rows; r++) {
    const int32_t* p = (const int32_t*)(m->data + (size_t)r * m->step_p[0]);
    for (int c = 0; c < m->cols; c++) s += (double)p[c];
  }
  ret[0] = s; ret[1] = ret[2] = ret[3] = 0;
  return ret;
}
static InputArray g_noarray = { 0, NULL, 0, 0 };
static MS const InputArray* s_noArray(void) { return &g_noarray; }

/* ------------------------------------------------------------------ C runtime stand-ins */
static MS void* s_memset(void* d, int c, size_t n) { return memset(d, c, n); }
static MS void* s_memcpy(void* d, const void* s, size_t n) { return memcpy(d, s, n); }
static MS void* s_memmove(void* d, const void* s, size_t n) { return memmove(d, s, n); }
static MS void* s_malloc(size_t n) { return malloc(n ? n : 1); }
static MS void* s_calloc(size_t a, size_t b) { return calloc(a ? a : 1, b ? b : 1); }
static MS void s_free(void* p) { free(p); }
static MS int s_callnewh(size_t n) { (void)n; return 0; }
static MS double s_sqrt(double x) { return sqrt(x); }
static MS void s_abort_named(const char* name) {
  fprintf(stderr, "gms_dll_host: the DLL called an import with no stand-in: %s\n", name);
  abort();
}
static MS void s_invalid_parameter(void) { fprintf(stderr, "gms_dll_host: _invalid_parameter_noinfo_noreturn\n"); abort(); }
static MS void s_xlength(const char* what) { fprintf(stderr, "gms_dll_host: std::length_error %s\n", what); abort(); }

static const struct { const char* name; void* fn; } g_known[] = {
  {"??0Mat@cv@@QEAA@XZ", (void*)s_mat_ctor},
  {"??1Mat@cv@@QEAA@XZ", (void*)s_mat_dtor},
  {"??0Mat@cv@@QEAA@AEBV01@AEBVRange@1@1@Z", (void*)s_mat_roi},
  {"?zeros@Mat@cv@@SA?AVMatExpr@2@HHH@Z", (void*)s_mat_zeros},
  {"?setTo@Mat@cv@@QEAAAEAV12@AEBV_InputArray@2@0@Z", (void*)s_mat_setTo},
  {"?sum@cv@@YA?AV?$Scalar_@N@1@AEBV_InputArray@1@@Z", (void*)s_cv_sum},
  {"?noArray@cv@@YAAEBV_InputOutputArray@1@XZ", (void*)s_noArray},
  {"memset", (void*)s_memset}, {"memcpy", (void*)s_memcpy}, {"memmove", (void*)s_memmove},
  {"malloc", (void*)s_malloc}, {"calloc", (void*)s_calloc}, {"free", (void*)s_free}, {"_callnewh", (void*)s_callnewh},
  {"sqrt", (void*)s_sqrt},
  {"_invalid_parameter_noinfo_noreturn", (void*)s_invalid_parameter},
  {"?_Xlength_error@std@@YAXPEBD@Z", (void*)s_xlength},
};

/* ------------------------------------------------------------------ PE32+ loader */
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

const char* gmsdll_error(void) { return g_err; }

int gmsdll_load(const char* path) {
  if (g_img) return 0;
  FILE* f = fopen(path, "rb");
  if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return -1; }
  fseek(f, 0, SEEK_END); long fsz = ftell(f); fseek(f, 0, SEEK_SET);
  uint8_t* file = (uint8_t*)malloc(fsz);
  if (fread(file, 1, fsz, f) != (size_t)fsz) { fclose(f); free(file); snprintf(g_err, sizeof g_err, "short read"); return -1; }
  fclose(f);
  if (fsz < 0x400 || file[0] != 'M' || file[1] != 'Z') { free(file); snprintf(g_err, sizeof g_err, "not a PE file"); return -1; }
  uint32_t pe = rd32(file + 0x3c);
  if (rd32(file + pe) != 0x4550 || rd16(file + pe + 4) != 0x8664 || rd16(file + pe + 24) != 0x20b) {
    free(file); snprintf(g_err, sizeof g_err, "not an x86-64 PE32+ image"); return -1;
  }
  int nsec = rd16(file + pe + 6);
  uint32_t optsz = rd16(file + pe + 20);
  const uint8_t* opt = file + pe + 24;
  uint64_t base = rd64(opt + 24);
  uint32_t image_size = rd32(opt + 56), hdr_size = rd32(opt + 60);
  if (base != IMAGE_BASE) { free(file); snprintf(g_err, sizeof g_err, "unexpected image base"); return -1; }
  uint8_t* img = (uint8_t*)mmap(NULL, image_size, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (img == MAP_FAILED) { free(file); snprintf(g_err, sizeof g_err, "mmap failed"); return -1; }
  memcpy(img, file, hdr_size);
  const uint8_t* sec = opt + optsz;
  for (int i = 0; i < nsec; i++, sec += 40) {
    uint32_t vsz = rd32(sec + 8), vaddr = rd32(sec + 12), rsz = rd32(sec + 16), roff = rd32(sec + 20);
    uint32_t n = rsz < vsz ? rsz : vsz;
    if ((uint64_t)vaddr + n > image_size || (uint64_t)roff + n > (uint64_t)fsz) { snprintf(g_err, sizeof g_err, "bad section"); return -1; }
    memcpy(img + vaddr, file + roff, n);
  }
  g_img = img; g_size = image_size;
  /* base relocations (data directory 5): only IMAGE_REL_BASED_DIR64 (10) and ABSOLUTE padding (0) occur */
  uint64_t delta = (uint64_t)img - IMAGE_BASE;
  uint32_t rel_rva = rd32(opt + 112 + 8 * 5), rel_sz = rd32(opt + 112 + 8 * 5 + 4);
  for (uint32_t o = 0; o + 8 <= rel_sz;) {
    uint32_t page = rd32(img + rel_rva + o), bsz = rd32(img + rel_rva + o + 4);
    if (bsz < 8) break;
    for (uint32_t k = 8; k + 2 <= bsz; k += 2) {
      uint16_t e = rd16(img + rel_rva + o + k);
      int type = e >> 12;
      if (type == 10) { uint64_t v = rd64(img + page + (e & 0xfff)) + delta; memcpy(img + page + (e & 0xfff), &v, 8); }
      else if (type != 0) { snprintf(g_err, sizeof g_err, "relocation type %d", type); return -1; }
    }
    o += bsz;
  }
  /* imports (data directory 1): known names -> stand-ins; everything else -> a generated thunk that names itself and aborts */
  uint32_t imp_rva = rd32(opt + 112 + 8 * 1);
  uint8_t* thunks = (uint8_t*)mmap(NULL, 1 << 16, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  size_t tpos = 0;
  for (uint32_t d = imp_rva; rd32(img + d + 12); d += 20) {
    uint32_t oft = rd32(img + d), ft = rd32(img + d + 16);
    uint32_t look = oft ? oft : ft;
    for (uint32_t k = 0;; k++) {
      uint64_t ent = rd64(img + look + 8 * k);
      if (!ent) break;
      const char* name = (ent >> 63) ? "(ordinal)" : (const char*)(img + (uint32_t)ent + 2);
      void* target = NULL;
      for (size_t q = 0; q < sizeof g_known / sizeof g_known[0]; q++)
        if (!strcmp(name, g_known[q].name)) target = g_known[q].fn;
      if (!target) {
        if (tpos + 32 > (1 << 16)) { snprintf(g_err, sizeof g_err, "thunk space"); return -1; }
        uint8_t* t = thunks + tpos; tpos += 32;
        uint64_t np = (uint64_t)name, fp = (uint64_t)(void*)s_abort_named;
        t[0] = 0x48; t[1] = 0xB9; memcpy(t + 2, &np, 8);      /* mov rcx, name   */
        t[10] = 0x48; t[11] = 0xB8; memcpy(t + 12, &fp, 8);   /* mov rax, abort  */
        t[20] = 0x48; t[21] = 0x83; t[22] = 0xE4; t[23] = 0xF0; /* and rsp,-16   */
        t[24] = 0xFF; t[25] = 0xD0;                           /* call rax        */
        target = t;
      }
      memcpy(img + ft + 8 * k, &target, 8);
    }
  }
  for (int i = 0; i < 16; i++) g_op_vtbl[i] = (void*)s_op_trap;
  g_op_vtbl[2] = (void*)s_op_assign;
  free(file);
  /* the DLL's dynamic initialiser for SCALE[2], SCALE[3] (1/sqrt 2, sqrt 2) — runs the DLL's own code */
  ((MS void (*)(void))va(0x1800010b0))();
  return 0;
}

/* ------------------------------------------------------------------ entry points used by the golden generator */
void gmsdll_tables(int32_t rot[72], double scale[5]) {
  memcpy(rot, va(0x18012f520), 72 * 4);
  memcpy(scale, va(0x1802c5008), 5 * 8);
}

/* MSVC std::vector<T>: {first, last, end}.  KeyPoint 28 B, DMatch 16 B. */
typedef struct Vec { uint8_t *first, *last, *end; } Vec;
typedef struct Size2i { int w, h; } Size2i;

/* leaf functions: int GMSMatcher::getGridIndexLeft(const Point2f&, int type), getGridIndexRight(const Point2f&) */
void gmsdll_grid_left(const float* pts_xy, int64_t n, int type, int32_t* out) {
  uint8_t obj[0x200]; memset(obj, 0, sizeof obj);
  *(int*)(obj + 0x50) = 20; *(int*)(obj + 0x54) = 20;
  MS int (*fn)(void*, const float*, int) = (MS int (*)(void*, const float*, int))va(0x180047bc0);
  for (int64_t i = 0; i < n; i++) out[i] = fn(obj, pts_xy + 2 * i, type);
}
void gmsdll_grid_right(const float* pts_xy, int64_t n, int wr, int hr, int32_t* out) {
  uint8_t obj[0x200]; memset(obj, 0, sizeof obj);
  *(int*)(obj + 0x50) = 20; *(int*)(obj + 0x54) = 20;
  *(int*)(obj + 0x58) = wr; *(int*)(obj + 0x5c) = hr;
  MS int (*fn)(void*, const float*) = (MS int (*)(void*, const float*))va(0x180047d60);
  for (int64_t i = 0; i < n; i++) out[i] = fn(obj, pts_xy + 2 * i);
}

typedef MS void (*match_gms_fn)(const Size2i*, const Size2i*, const Vec*, const Vec*, const Vec*, Vec*, uint8_t, uint8_t, double);

/* The export itself.  keypoints: n x 28-byte cv::KeyPoint; matches: n x 16-byte cv::DMatch.
 * out must hold n_matches DMatch; returns the number written (matchesGMS.size()). */
int64_t gmsdll_match_gms(int w1, int h1, int w2, int h2, const uint8_t* kp1, int64_t n1, const uint8_t* kp2, int64_t n2,
                         const uint8_t* matches, int64_t n, int with_rotation, int with_scale, double factor, uint8_t* out) {
  Size2i s1 = {w1, h1}, s2 = {w2, h2};
  Vec v1 = {(uint8_t*)kp1, (uint8_t*)kp1 + 28 * n1, (uint8_t*)kp1 + 28 * n1};
  Vec v2 = {(uint8_t*)kp2, (uint8_t*)kp2 + 28 * n2, (uint8_t*)kp2 + 28 * n2};
  Vec vm = {(uint8_t*)matches, (uint8_t*)matches + 16 * n, (uint8_t*)matches + 16 * n};
  /* output vector with capacity n so that push_back never reallocates through the MSVC allocator */
  Vec vo = {out, out, out + 16 * n};
  ((match_gms_fn)va(0x180048280))(&s1, &s2, &v1, &v2, &vm, &vo, (uint8_t)with_rotation, (uint8_t)with_scale, factor);
  return (vo.last - vo.first) / 16;
}

/* One hypothesis at a time: ctor, setScale(s), run(rot) -> inlier count and the object's mvbInlierMask (+0x110,
 * MSVC vector<bool> = vector<uint32> + size).  counts[s*8 + (rot-1)]; masks (optional) [40][n] bytes. */
int gmsdll_hypotheses(int w1, int h1, int w2, int h2, const uint8_t* kp1, int64_t n1, const uint8_t* kp2, int64_t n2,
                      const uint8_t* matches, int64_t n, double factor, int32_t* counts, uint8_t* masks) {
  Size2i s1 = {w1, h1}, s2 = {w2, h2};
  Vec v1 = {(uint8_t*)kp1, (uint8_t*)kp1 + 28 * n1, (uint8_t*)kp1 + 28 * n1};
  Vec v2 = {(uint8_t*)kp2, (uint8_t*)kp2 + 28 * n2, (uint8_t*)kp2 + 28 * n2};
  Vec vm = {(uint8_t*)matches, (uint8_t*)matches + 16 * n, (uint8_t*)matches + 16 * n};
  uint8_t* obj = (uint8_t*)calloc(1, 0x400);
  /* GMSMatcher(kp1, size1, kp2, size2, matches, thresholdFactor): args as matchGMS passes them (0x1800482a4-0x1800482dc) */
  ((MS void* (*)(void*, const Vec*, const Size2i*, const Vec*, const Size2i*, const Vec*, double))va(0x180046900))(
      obj, &v1, &s1, &v2, &s2, &vm, factor);
  for (int s = 0; s < 5; s++) {
    ((MS void (*)(void*, int))va(0x180048c10))(obj, s);
    for (int r = 1; r <= 8; r++) {
      int c = ((MS int (*)(void*, int))va(0x180048630))(obj, r);
      counts[s * 8 + r - 1] = c;
      if (masks) {
        const uint32_t* bits = *(const uint32_t**)(obj + 0x110);
        uint64_t sz = *(uint64_t*)(obj + 0x128);
        if ((int64_t)sz != n) { free(obj); snprintf(g_err, sizeof g_err, "mask size %llu != %lld", (unsigned long long)sz, (long long)n); return -1; }
        uint8_t* m = masks + (size_t)(s * 8 + r - 1) * n;
        for (int64_t i = 0; i < n; i++) m[i] = (bits[i >> 5] >> (i & 31)) & 1;
      }
    }
  }
  ((MS void (*)(void*))va(0x180046d20))(obj);
  free(obj);
  return 0;
}
