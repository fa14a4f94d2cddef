// flashe_stats.cu — per-layer statistics around decode, BIT-EXACT with numpy.
//
//   QuantizingClient.unnormalize (secureprotol/jzf_quantize.py:549-564), per layer:
//       w += past_mean[layer];  past_mean[layer] = np.mean(w);  past_std[layer] = np.std(w)
//   The standard deviation of one round defines the clipping range alpha of the next
//   (jzf_quantize.py:403-413), and alpha is part of every ciphertext: a last-bit difference in std can move
//   float32(alpha) and with it the quantised values of a GPU party against a numpy party.  The sums are
//   therefore evaluated in numpy's own order:
//
//   FLASHE_SUM_PAIRWISE    float64 ndarrays (what the batched mode holds after unbatch -> unquantize):
//       np.add.reduce = 0.0 + pairwise_sum(a, n) with numpy's fixed recursion (loops_utils.h.src,
//       @TYPE@_pairwise_sum, unchanged since numpy 1.9; the reference pins 1.17.2):
//           n < 8        sequential from 0.0
//           n <= 128     eight accumulators r[j] = a[j] + a[8+j] + a[16+j] + ..., combined as
//                        ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), then the n % 8 tail added one by one
//           n > 128      split at n2 = (n/2) - (n/2) % 8 and add the two halves' sums
//       mean = sum / n;  std = sqrt( (0.0 + pairwise_sum((a - mean)^2)) / n )      (numpy/core/_methods.py)
//   FLASHE_SUM_SEQUENTIAL  object arrays of Python floats (what the un-batched mode holds: unquantize of an
//       object array yields Python floats): no identity, a[0] + a[1] + ... left to right; same mean / std
//       formulas.  Inherently serial per layer (one lane adds, the warp stages the data).
//
// The pairwise tree depends only on the sizes, so it is cut into three levels that are each enumerated
// where that is cheap: the HOST walks the recursion down to nodes of at most GROUP_ELEMS elements ("groups")
// and MID_ELEMS elements ("mid nodes") and lists the <= 128-element leaves below a group once per distinct
// group size; one WARP per group evaluates its leaves eight lanes per leaf exactly as numpy's unrolled loop
// does and folds them up the group's tree; one thread per mid node combines its groups, one thread per layer
// its mid nodes.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "flashe_internal.h"

#define GROUP_ELEMS 8192u       // a group is a tree node with <= 8192 elements whose parent has more
#define MID_ELEMS 262144u
#define SG_WARPS 8

struct StatGroup { uint64_t begin; uint32_t n; uint32_t seg; uint32_t shape; uint32_t nleaf; };   // shape: first leaf descriptor
struct StatMid { uint64_t n; uint32_t first_group; uint32_t seg; };
struct StatSegD { uint64_t begin, n; uint32_t first_mid, n_mid; double shift; };

static __host__ __device__ __forceinline__ uint64_t pw_left(uint64_t n) { uint64_t h = n >> 1; return h - (h & 7ull); }

// combine the sums of consecutive sub-nodes of a node with `n` elements: sub-nodes are the nodes of the
// recursion with at most `cut` elements; *cur walks their sums in depth-first order
__device__ double pw_combine(uint64_t n, uint64_t cut, const double* sums, uint32_t* cur) {
    if (n <= cut) return sums[(*cur)++];
    const uint64_t n2 = pw_left(n);
    const double a = pw_combine(n2, cut, sums, cur);
    const double b = pw_combine(n - n2, cut, sums, cur);
    return __dadd_rn(a, b);
}

// Leaf descriptors of a group shape (the tree below a group depends only on its size, so groups of equal size
// share one list, built by the host): bits 0-12 offset of the leaf in the group, 13-19 its size - 1, 20-27 its
// heap index in the group's tree (root 1, children 2i and 2i+1; depth <= 7 because sizes halve down to <= 128).
#define LEAF_OFF(d) ((d) & 0x1fffu)
#define LEAF_N(d) ((((d) >> 13) & 0x7fu) + 1u)
#define LEAF_HEAP(d) (((d) >> 20) & 0xffu)

// PASS 0: v = w + shift (stored to w_out when given), group sum of v.
// PASS 1: group sum of (x - mean)^2, x = w_out, or w + shift when no w_out was written.
// One warp per group.  Leaves go four at a time, eight lanes each: lane j of a leaf owns numpy's accumulator
// r[j] (all of its <= 16 loads are issued before the adds).  Leaf sums land in the group's heap-indexed node
// array in shared memory, which is then folded level by level (children 2i, 2i+1 -> i), 32 nodes per step.
template <int PASS>
__global__ void __launch_bounds__(SG_WARPS * 32)
k_stats_groups(const double* __restrict__ w, double* __restrict__ w_out, const StatGroup* __restrict__ groups, uint32_t ngroups,
               const StatSegD* __restrict__ segs, const uint32_t* __restrict__ shapes, int add_shift,
               const double* __restrict__ stats, double* __restrict__ gsum) {
    __shared__ double node[SG_WARPS][256];
    __shared__ uint8_t present[SG_WARPS][256];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t sub = lane >> 3, j = lane & 7u;            // four 8-lane groups per warp
    const uint32_t gmask = 0xffu << (8u * sub);
    for (uint32_t g = blockIdx.x * SG_WARPS + warp; g < ngroups; g += gridDim.x * SG_WARPS) {
        const StatGroup gr = groups[g];
        const double shift = segs[gr.seg].shift;
        const double mean = PASS == 1 ? stats[2 * gr.seg] : 0.0;
        const double* src = w + gr.begin;
        double* dst = (PASS == 0 && w_out) ? w_out + gr.begin : nullptr;
        reinterpret_cast<uint64_t*>(present[warp])[lane] = 0ull;
        __syncwarp();
        auto val = [&](uint32_t idx) -> double {               // element idx of the group, as the pass sees it
            if (PASS == 0) {
                const double v = __dadd_rn(src[idx], shift);
                if (dst) dst[idx] = v;
                return v;
            } else {
                const double x = add_shift ? __dadd_rn(src[idx], shift) : src[idx];
                const double d = __dsub_rn(x, mean);
                return __dmul_rn(d, d);
            }
        };
        for (uint32_t l0 = 0; l0 < gr.nleaf; l0 += 4u) {
            const uint32_t l = l0 + sub;
            if (l < gr.nleaf) {                                 // uniform over the 8 lanes of a leaf
                const uint32_t desc = __ldg(shapes + gr.shape + l);
                const uint32_t off = LEAF_OFF(desc), n = LEAF_N(desc);
                double r;
                if (n < 8u) {                                   // (only a whole layer can be this small)
                    r = 0.0;
                    for (uint32_t i = 0; i < n; ++i) r = __dadd_rn(r, val(off + i));
                } else {
                    const uint32_t n8 = n - (n & 7u);
                    double v[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = (uint32_t)(8 * k) < n8 ? val(off + 8u * k + j) : 0.0;
                    r = v[0];
#pragma unroll
                    for (int k = 1; k < 16; ++k) if ((uint32_t)(8 * k) < n8) r = __dadd_rn(r, v[k]);
                    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));          // r0+r1 | r2+r3 | r4+r5 | r6+r7
                    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));          // (r0+r1)+(r2+r3) | (r4+r5)+(r6+r7)
                    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
                    for (uint32_t i = n8; i < n; ++i) r = __dadd_rn(r, val(off + i));   // the n % 8 tail, one by one
                }
                if (j == 0u) { node[warp][LEAF_HEAP(desc)] = r; present[warp][LEAF_HEAP(desc)] = 1; }
            }
        }
        __syncwarp();
        // fold the tree bottom-up: every internal node has both children
#pragma unroll 1
        for (int lvl = 7; lvl >= 1; --lvl) {
            const uint32_t first = 1u << lvl, cnt = 1u << (lvl - 1);       // pairs at this level
            for (uint32_t p = lane; p < cnt; p += 32u) {
                const uint32_t a = first + 2u * p;
                if (present[warp][a]) {
                    node[warp][a >> 1] = __dadd_rn(node[warp][a], node[warp][a + 1u]);
                    present[warp][a >> 1] = 1;
                }
            }
            __syncwarp();
        }
        if (lane == 0) gsum[g] = node[warp][1];
        __syncwarp();
    }
}

__global__ void k_stats_mids(const StatMid* __restrict__ mids, uint32_t nmid, const double* __restrict__ gsum, double* __restrict__ msum) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmid) return;
    uint32_t cur = 0;
    msum[m] = pw_combine(mids[m].n, GROUP_ELEMS, gsum + mids[m].first_group, &cur);
}

template <int PASS>
__global__ void k_stats_top(const StatSegD* __restrict__ segs, int nseg, const double* __restrict__ msum, double* __restrict__ stats) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const StatSegD sg = segs[s];
    if (sg.n == 0) {                        // np.mean / np.std of an empty array: nan (with a RuntimeWarning)
        stats[2 * s + PASS] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    uint32_t cur = 0;
    const double tot = __dadd_rn(0.0, pw_combine(sg.n, MID_ELEMS, msum + sg.first_mid, &cur));   // identity + pairwise sum
    const double q = __ddiv_rn(tot, (double)sg.n);
    stats[2 * s + PASS] = PASS == 0 ? q : __dsqrt_rn(q);
}

// Object-array order: a[0] + a[1] + ... left to right.  One warp per layer: the warp stages 256 elements at a
// time in shared memory (coalesced), lane 0 adds them in order.
#define SEQ_CHUNK 256
__global__ void __launch_bounds__(32)
k_stats_sequential(const double* __restrict__ w, double* __restrict__ w_out, const StatSegD* __restrict__ segs, int nseg,
                   double* __restrict__ stats) {
    __shared__ double buf[2][SEQ_CHUNK];
    const int s = blockIdx.x;
    if (s >= nseg) return;
    const StatSegD sg = segs[s];
    const uint32_t lane = threadIdx.x;
    if (sg.n == 0) { if (lane == 0) { stats[2 * s] = stats[2 * s + 1] = __longlong_as_double(0x7ff8000000000000ll); } return; }
    double mean = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        double acc = 0.0;
        const double* src = (pass == 1 && w_out) ? w_out + sg.begin : w + sg.begin;
        const bool shift_here = pass == 0 || !w_out;
        auto stage = [&](uint64_t c0, int b) {
            for (uint32_t i = lane; i < SEQ_CHUNK; i += 32u) {
                const uint64_t e = c0 + i;
                if (e < sg.n) {
                    double v = shift_here ? __dadd_rn(src[e], sg.shift) : src[e];
                    if (pass == 0) { if (w_out) w_out[sg.begin + e] = v; }
                    else { const double d = __dsub_rn(v, mean); v = __dmul_rn(d, d); }
                    buf[b][i] = v;
                }
            }
        };
        stage(0, 0);
        __syncwarp();
        int b = 0;
        for (uint64_t c0 = 0; c0 < sg.n; c0 += SEQ_CHUNK, b ^= 1) {
            if (c0 + SEQ_CHUNK < sg.n) stage(c0 + SEQ_CHUNK, b ^ 1);           // next chunk in flight while lane 0 adds
            if (lane == 0) {
                const uint32_t m = (uint32_t)(sg.n - c0 < SEQ_CHUNK ? sg.n - c0 : SEQ_CHUNK);
                uint32_t i = 0;
                if (c0 == 0) { acc = buf[b][0]; i = 1; }                        // no identity: the sum starts at a[0]
                for (; i < m; ++i) acc = __dadd_rn(acc, buf[b][i]);
            }
            __syncwarp();
        }
        if (pass == 0) {
            mean = __ddiv_rn(__shfl_sync(0xffffffffu, acc, 0), (double)sg.n);
            if (lane == 0) stats[2 * s] = mean;
        } else if (lane == 0) {
            stats[2 * s + 1] = __dsqrt_rn(__ddiv_rn(acc, (double)sg.n));
        }
        __syncwarp();
    }
}

// host: walk the recursion of one layer down to the groups, recording the mid nodes on the way
static void shape_leaves(uint32_t off, uint32_t n, uint32_t heap, std::vector<uint32_t>& out) {
    if (n <= 128u) { out.push_back(off | ((n - 1u) << 13) | (heap << 20)); return; }
    const uint32_t n2 = (uint32_t)pw_left(n);
    shape_leaves(off, n2, 2u * heap, out);
    shape_leaves(off + n2, n - n2, 2u * heap + 1u, out);
}

static void plan_node(uint64_t begin, uint64_t n, uint32_t seg, bool in_mid, std::vector<StatGroup>& groups, std::vector<StatMid>& mids) {
    if (!in_mid && n <= MID_ELEMS) {
        StatMid m; m.n = n; m.first_group = (uint32_t)groups.size(); m.seg = seg;
        mids.push_back(m);
        in_mid = true;
    }
    if (n <= GROUP_ELEMS) {
        StatGroup g; g.begin = begin; g.n = (uint32_t)n; g.seg = seg; g.shape = 0; g.nleaf = 0;
        groups.push_back(g);
        return;
    }
    const uint64_t n2 = pw_left(n);
    plan_node(begin, n2, seg, in_mid, groups, mids);
    plan_node(begin + n2, n - n2, seg, in_mid, groups, mids);
}

// A layout's plan, resident on its device.  Up to STAT_PLANS layouts are kept (least recently used goes first);
// a plan in use by a call stays alive through the shared_ptr even if it is evicted meanwhile, and its device
// tables are only read by kernels that were enqueued before the free (cudaFree synchronises the device).
struct StatPlan {
    int device = 0, order = 0;
    std::vector<uint64_t> seg_end;
    std::vector<StatSegD> segs;            // shift = 0: filled per call
    uint32_t n_groups = 0, n_mids = 0;
    StatGroup* d_groups = nullptr; StatMid* d_mids = nullptr; uint32_t* d_shapes = nullptr;
    uint64_t stamp = 0;
    ~StatPlan() {
        int prev = -1;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) cudaSetDevice(device);
        if (d_groups) cudaFree(d_groups);
        if (d_mids) cudaFree(d_mids);
        if (d_shapes) cudaFree(d_shapes);
        if (prev >= 0 && prev != device) cudaSetDevice(prev);
    }
};
#define STAT_PLANS 8
static std::mutex g_plan_mu;
static std::vector<std::shared_ptr<StatPlan>> g_plans;
static uint64_t g_plan_clock = 0;

static int stats_plan(int device, const uint64_t* seg_end, int nseg, int order, std::shared_ptr<StatPlan>* out) {
    std::lock_guard<std::mutex> lock(g_plan_mu);
    for (auto& p : g_plans)
        if (p->device == device && p->order == order && p->seg_end.size() == (size_t)nseg && memcmp(p->seg_end.data(), seg_end, 8 * (size_t)nseg) == 0) {
            p->stamp = ++g_plan_clock; *out = p; return FLASHE_OK;
        }
    auto p = std::make_shared<StatPlan>();
    p->device = device; p->order = order; p->seg_end.assign(seg_end, seg_end + nseg);
    p->segs.resize((size_t)nseg);
    std::vector<StatGroup> groups;
    std::vector<StatMid> mids;
    uint64_t prev = 0;
    for (int s = 0; s < nseg; ++s) {
        StatSegD& sg = p->segs[(size_t)s];
        sg.begin = prev; sg.n = seg_end[s] - prev; sg.shift = 0.0;
        sg.first_mid = (uint32_t)mids.size();
        if (order == FLASHE_SUM_PAIRWISE && sg.n) plan_node(prev, sg.n, (uint32_t)s, false, groups, mids);
        sg.n_mid = (uint32_t)mids.size() - sg.first_mid;
        prev = seg_end[s];
    }
    if (groups.size() > 0xfffffff0ull) return flashe_fail(FLASHE_EUNSUPPORTED, "vector too long for one statistics call");
    // one leaf list per distinct group size
    std::vector<uint32_t> shapes;
    {
        std::unordered_map<uint32_t, std::pair<uint32_t, uint32_t>> seen;
        for (auto& g : groups) {
            auto it = seen.find(g.n);
            if (it == seen.end()) {
                const uint32_t first = (uint32_t)shapes.size();
                shape_leaves(0u, g.n, 1u, shapes);
                it = seen.emplace(g.n, std::make_pair(first, (uint32_t)shapes.size() - first)).first;
            }
            g.shape = it->second.first; g.nleaf = it->second.second;
        }
    }
    p->n_groups = (uint32_t)groups.size(); p->n_mids = (uint32_t)mids.size();
    cudaError_t e = cudaSuccess;
    if (!groups.empty()) { e = cudaMalloc((void**)&p->d_groups, sizeof(StatGroup) * groups.size()); if (e == cudaSuccess) e = cudaMemcpy(p->d_groups, groups.data(), sizeof(StatGroup) * groups.size(), cudaMemcpyHostToDevice); }
    if (e == cudaSuccess && !mids.empty()) { e = cudaMalloc((void**)&p->d_mids, sizeof(StatMid) * mids.size()); if (e == cudaSuccess) e = cudaMemcpy(p->d_mids, mids.data(), sizeof(StatMid) * mids.size(), cudaMemcpyHostToDevice); }
    if (e == cudaSuccess && !shapes.empty()) { e = cudaMalloc((void**)&p->d_shapes, 4 * shapes.size()); if (e == cudaSuccess) e = cudaMemcpy(p->d_shapes, shapes.data(), 4 * shapes.size(), cudaMemcpyHostToDevice); }
    if (e != cudaSuccess) { cudaGetLastError(); return flashe_fail(FLASHE_ECUDA, std::string("flashe_segment_stats (plan): ") + cudaGetErrorString(e)); }
    p->stamp = ++g_plan_clock;
    if (g_plans.size() >= STAT_PLANS) {
        size_t victim = 0;
        for (size_t i = 1; i < g_plans.size(); ++i) if (g_plans[i]->stamp < g_plans[victim]->stamp) victim = i;
        g_plans.erase(g_plans.begin() + (long)victim);
    }
    g_plans.push_back(p);
    *out = p;
    return FLASHE_OK;
}

extern "C" int flashe_segment_stats(flashe_ctx* ctx, const double* w, double* w_out, uint64_t total, const uint64_t* seg_end,
                                    const double* shift, int nseg, int order, double* stats_out, void* stream) {
    flashe_ctx_info info;
    { int rc = flashe_ctx_get_info(ctx, &info); if (rc) return rc; }
    FlasheDeviceGuard guard(info.device);
    if (!guard.ok) return flashe_fail(FLASHE_ECUDA, "cudaSetDevice failed");
    cudaStream_t cs = (cudaStream_t)stream;
    if (nseg < 1 || !seg_end) return flashe_fail(FLASHE_EINVAL, "need nseg >= 1 and seg_end");
    if (seg_end[nseg - 1] != total) return flashe_fail(FLASHE_EINVAL, "seg_end[nseg-1] must equal total");
    if (!stats_out) return flashe_fail(FLASHE_EINVAL, "stats_out is NULL");
    if (total && !w) return flashe_fail(FLASHE_EINVAL, "w is NULL");
    if (order != FLASHE_SUM_PAIRWISE && order != FLASHE_SUM_SEQUENTIAL) return flashe_fail(FLASHE_EINVAL, "unknown summation order");
    // The plan (groups, mid nodes, leaf shapes) depends only on the layer sizes: it is built once per layout and kept
    // on the device (a model's layout is the same every round); only the per-layer shifts travel with each call.
    std::vector<StatSegD> segs((size_t)nseg);
    uint64_t prev = 0;
    for (int s = 0; s < nseg; ++s) {
        if (seg_end[s] < prev) return flashe_fail(FLASHE_EINVAL, "seg_end must be ascending");
        prev = seg_end[s];
    }
    std::shared_ptr<StatPlan> plan;
    { int rc = stats_plan(info.device, seg_end, nseg, order, &plan); if (rc) return rc; }
    for (int s = 0; s < nseg; ++s) { segs[s] = plan->segs[(size_t)s]; segs[s].shift = shift ? shift[s] : 0.0; }
    const uint32_t n_groups = plan->n_groups, n_mids = plan->n_mids;
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_seg = pad(sizeof(StatSegD) * segs.size()), b_gs = pad(8 * (size_t)n_groups), b_ms = pad(8 * (size_t)n_mids);
    uint8_t* ws = nullptr;
    FLASHE_CUDA_TRY(cudaMallocAsync((void**)&ws, b_seg + b_gs + b_ms + 256, cs));
    StatSegD* dseg = reinterpret_cast<StatSegD*>(ws);
    double* gsum = reinterpret_cast<double*>(ws + b_seg);
    double* msum = reinterpret_cast<double*>(ws + b_seg + b_gs);
    const StatGroup* dgrp = plan->d_groups;
    const StatMid* dmid = plan->d_mids;
    const uint32_t* dshape = plan->d_shapes;
    // (a pageable source has been staged by the time cudaMemcpyAsync returns: `segs` may go out of scope)
    cudaError_t e = cudaMemcpyAsync(dseg, segs.data(), sizeof(StatSegD) * segs.size(), cudaMemcpyHostToDevice, cs);
    int launches = 0;
    if (e == cudaSuccess && order == FLASHE_SUM_SEQUENTIAL) {
        k_stats_sequential<<<nseg, 32, 0, cs>>>(w, w_out, dseg, nseg, stats_out);
        launches = 1;
        e = cudaGetLastError();
    } else if (e == cudaSuccess) {
        const uint32_t ng = n_groups, nm = n_mids;
        const uint64_t cap = (uint64_t)info.num_sms * 8;
        const uint64_t want = (ng + SG_WARPS - 1) / SG_WARPS;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        const unsigned gm = nm ? (nm + 127) / 128 : 1, gt = (unsigned)((nseg + 127) / 128);
        if (ng) k_stats_groups<0><<<grid, SG_WARPS * 32, 0, cs>>>(w, w_out, dgrp, ng, dseg, dshape, 0, stats_out, gsum);
        if (nm) k_stats_mids<<<gm, 128, 0, cs>>>(dmid, nm, gsum, msum);
        k_stats_top<0><<<gt, 128, 0, cs>>>(dseg, nseg, msum, stats_out);
        if (ng) k_stats_groups<1><<<grid, SG_WARPS * 32, 0, cs>>>(w_out ? w_out : w, nullptr, dgrp, ng, dseg, dshape, w_out ? 0 : 1, stats_out, gsum);
        if (nm) k_stats_mids<<<gm, 128, 0, cs>>>(dmid, nm, gsum, msum);
        k_stats_top<1><<<gt, 128, 0, cs>>>(dseg, nseg, msum, stats_out);
        launches = 2 + (ng ? 2 : 0) + (nm ? 2 : 0);
        e = cudaGetLastError();
    }
    flashe_count_launches(launches);
    cudaFreeAsync(ws, cs);
    if (e != cudaSuccess) return flashe_fail(FLASHE_ECUDA, std::string("flashe_segment_stats: ") + cudaGetErrorString(e));
    return FLASHE_OK;
}
