// One frame = this launch sequence.  Nothing in it depends on the camera, the output buffers or the
// element counts: those live in device memory (SageFrameParams, counters), every grid is fixed, so
// the sequence is captured once per frame context into a CUDA graph and replayed.
//
//   caller stream cs ──fork──► front stream fs (highest priority):
//        memset(counters, ranges) → K1a project ─┬─► depth sort (3 passes x hist/scan/scatter) → K2 emit
//                                                 │        → by-bin sort (2 passes; the last writes K4 ranges) → order ─┐
//                                                 └─► side stream ss: K1b SH colour ────────────────────────────┤
//   cs ◄──join── K5 composite ◄─────────────────────────────────────────────────────────────────────────────────┘
//
// The front half is a chain of small latency-bound kernels; the composite is one long issue-bound
// kernel.  On a high-priority stream the front half of frame i+1 is dispatched ahead of the queued
// composite CTAs of frame i (another context), so consecutive frames overlap.
#include <cstdlib>

#include "api_common.cuh"

// tuning knobs (experiments): items per thread of the two sorts' tiles (8 or 16)
static int env_ipt(const char* name, int dflt)
{
    const char* v = getenv(name);
    if (!v) return dflt;
    const int k = atoi(v);
    return (k == 8 || k == 16) ? k : dflt;
}

static int depth_passes_of(int depth_bits, int* digit_bits)
{
    int rb = 0, np = 0;
    sage_sort_shape(depth_bits, &rb, &np);
    *digit_bits = rb;
    return np;
}

int sage_frame_enqueue(const SageSceneDev& sc, const SageFrame& fr, const SageFrameShape& shape,
                       const uint32_t* n_dev, cudaStream_t cs, cudaStream_t fs, cudaStream_t ss,
                       cudaEvent_t* fork_join /* 4 events */, cudaEvent_t* ev, int n_bins)
{
    int launches = 0;
    const bool serial = ev != nullptr;              // stage timing: everything on one stream, SH colour in line
    if (serial) { fs = cs; ss = cs; }
#define MARK(k) do { if (ev) cudaEventRecord(ev[k], cs); } while (0)
    if (!serial) {
        cudaEventRecord(fork_join[0], cs);
        cudaStreamWaitEvent(fs, fork_join[0], 0);
    }
    cudaMemsetAsync(fr.counters, 0, sizeof(uint32_t) * SAGE_ZERO_WORDS, fs);
    if (n_bins > 0) cudaMemsetAsync(fr.ranges, 0, sizeof(uint2) * (size_t)n_bins, fs);
    MARK(0);
    if (sc.n > 0 && n_bins > 0) {
        sage_launch_project(sc, fr, fs);
        launches += 1;
        MARK(1);
        if (!serial) {
            cudaEventRecord(fork_join[1], fs);
            cudaStreamWaitEvent(ss, fork_join[1], 0);
            sage_launch_sh_color(sc, fr, ss);
            cudaEventRecord(fork_join[2], ss);
            launches += 1;
        }
        SageSortPlan dp{};
        dp.keys_src = fr.dkey_src;
        dp.keys[0] = fr.dkey[0]; dp.keys[1] = fr.dkey[1];
        dp.vals[0] = fr.dval[0]; dp.vals[1] = fr.dval[1];
        dp.n_first = n_dev; dp.n_rest = fr.counters + CNT_VISIBLE;
        dp.lo_bit = 0; dp.n_bits = shape.depth_bits;
        dp.passes = depth_passes_of(shape.depth_bits, &dp.digit_bits);
        static const int depth_ipt = env_ipt("SAGE_DEPTH_IPT", 8);
        dp.vals_are_iota = true; dp.compact_first = true; dp.items_per_thread = depth_ipt;
        dp.H = fr.sort_h; dp.ld = fr.sort_ld; dp.totals = fr.sort_totals; dp.ranges = nullptr;
        const int res = sage_radix_sort(dp, &launches, fs);
        if (res < 0) return -1;
        MARK(2);
        if (serial) { sage_launch_sh_color(sc, fr, cs); launches += 1; }
        MARK(3);
        sage_launch_emit(fr, fr.dval[res], fs);
        launches += 1;
        MARK(4);
        SageSortPlan tp{};
        tp.keys_src = fr.ikey[0];
        tp.keys[0] = fr.ikey[0]; tp.keys[1] = fr.ikey[1];
        tp.vals[0] = fr.ival[0]; tp.vals[1] = fr.ival[1];
        tp.n_first = tp.n_rest = fr.counters + CNT_INSTANCES;
        // digits just wide enough for the tile ids (1080p: 8160 tiles = 13 bits -> two passes of 7 bits, not 8:
        // half the warp-private counters to clear and combine, one ballot fewer per item)
        tp.lo_bit = 0; tp.passes = shape.tile_passes; tp.digit_bits = shape.tile_digit_bits;
        tp.n_bits = tp.digit_bits * tp.passes;
        static const int tile_ipt = env_ipt("SAGE_TILE_IPT", 8);
        tp.vals_are_iota = false; tp.compact_first = false; tp.items_per_thread = tile_ipt;
        tp.H = fr.sort_h; tp.ld = fr.sort_ld; tp.totals = fr.sort_totals; tp.ranges = fr.ranges;
        const int ires = sage_radix_sort(tp, &launches, fs);
        if (ires < 0) return -1;
        sage_launch_bin_order(fr, fs);
        launches += 1;
        MARK(5);
        if (!serial) {
            cudaEventRecord(fork_join[3], fs);
            cudaStreamWaitEvent(cs, fork_join[3], 0);
            cudaStreamWaitEvent(cs, fork_join[2], 0);
        }
        sage_launch_composite(fr, fr.ival[ires], n_bins, shape.whole_frame != 0, cs);
        launches += 1;
        MARK(6);
    } else {
        for (int k = 1; k <= 5; ++k) MARK(k);
        if (n_bins > 0) {
            sage_launch_bin_order(fr, fs);               // D = 0: every bin is empty
            launches += 1;
        }
        if (!serial) {
            cudaEventRecord(fork_join[3], fs);
            cudaStreamWaitEvent(cs, fork_join[3], 0);
        }
        if (n_bins > 0) { sage_launch_composite(fr, fr.ival[0], n_bins, shape.whole_frame != 0, cs); launches += 1; }
        MARK(6);
    }
#undef MARK
    return launches;
}
