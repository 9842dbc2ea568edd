// hamming_tc.cuh — interface of the tcgen05 (int8 tensor-core) Hamming kernel, see hamming_tc.cu.
#pragma once
#include "common.cuh"

namespace sfmgms {

struct TcState {
    void* d_ops = nullptr;      // unpacked +-1 int8 operands of the registered image set / ad-hoc pair
    size_t ops_cap = 0;
    void* d_work = nullptr;     // work-unit table
    size_t work_cap = 0;
    void* h_work = nullptr;     // pinned
    size_t h_work_cap = 0;
    void* d_cnt = nullptr;      // fp4 kernel, fused tie resolution: arrival counters of split units (all zero between launches)
    size_t cnt_cap = 0;
    size_t work_used = 0;       // arena cursor (bytes) into h_work/d_work; reset by tc_reset_arena() after a sync
    bool set_valid = false;     // d_ops holds the unpacked form of [ops_src, ops_src + 32*ops_rows)
    bool cache_enabled = true;  // keep the unpacked operands across calls until tc_invalidate()
    const uint8_t* ops_src = nullptr;
    long long ops_rows = 0;
    int ops_row_bytes = 0;      // 256 (int8 kernel) or 128 (fp4 kernel)
    // A registered image set pins the operand span to the WHOLE set, so every launch over it (any chunk of any pair
    // list) finds the same unpacked array: each image is unpacked once per set, not once per launch.
    const uint8_t* span_lo = nullptr;
    long long span_rows = 0;
};

inline void tc_pin_span(TcState& s, const uint8_t* lo, long long rows) { s.span_lo = lo; s.span_rows = rows; }
// widen [lo, hi) to the pinned span when it lies inside it
inline void tc_apply_span(const TcState& s, const uint8_t*& lo, const uint8_t*& hi) {
    if (s.cache_enabled && s.span_lo && lo >= s.span_lo && hi <= s.span_lo + (size_t)s.span_rows * 32) {
        lo = s.span_lo;
        hi = s.span_lo + (size_t)s.span_rows * 32;
    }
}

bool tc_available();
const char* tc_last_error();
void tc_invalidate(TcState& s);
void tc_release(TcState& s);
// The work-unit tables of all launches since the last stream synchronisation live side by side in a pinned
// arena (the CPU must not overwrite a table whose H2D copy is still queued).  Call after synchronising.
void tc_reset_arena(TcState& s);
// returns the number of kernel launches, or -1 on error
int launch_hamming_tc(TcState& s, const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int sm_count,
                      cudaStream_t st);

// FP4 (tcgen05 kind::mxf4 block-scaled) variant, hamming_fp4.cu — same contract as launch_hamming_tc
const char* fp4_last_error();
// By default the kernel settles the tie rule itself (two otherwise idle warps per CTA, one unit behind the epilogue);
// SFMGMS_FP4_FUSED_RESOLVE=0 brings back the separate hamming_resolve_kernel (fp4_fused_resolve() tells which).
// resolve_st / resolve_ev (optional, separate kernel only): run it on a second stream, ordered after the tensor-core kernel
// through the event; the caller joins that stream before anything reads the keys.
bool fp4_fused_resolve();
int launch_hamming_fp4(TcState& s, const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int sm_count,
                       cudaStream_t st, cudaStream_t resolve_st = nullptr, cudaEvent_t resolve_ev = nullptr);

}  // namespace sfmgms
