// smcb_scan.cuh -- single-pass inclusive prefix sum of non-negative fp64 values
// (the CDF walked by inverse_cdf, particles/resampling.py:484-509).
//
// Decoupled look-back (Merrill & Garland 2016) with two changes that make the
// result a pure function of the input:
//   * tiles publish only their AGGREGATE; prefixes are rebuilt from aggregates with
//     a fixed association: groups of 32 consecutive tiles are summed by a warp scan,
//     group prefixes C_g are chained sequentially (C_{g+1} = C_g + S_g).  A tile may
//     pick up an already-published C_k as a shortcut, but the value it would have
//     computed itself is bit-identical, so timing never changes a bit of the output.
//   * every level (thread, warp, block, group) clamps its values into the interval
//     spanned by its own base and the base of its successor, which makes the output
//     non-decreasing BY CONSTRUCTION.  np.searchsorted on it is then well defined and
//     the search kernel can be held to it bit-exactly.
#pragma once
#include "smcb_common.cuh"

namespace smcb {

constexpr int kScanItems = 8;                       // fp64 values per thread
constexpr int kScanTile = kBlock * kScanItems;      // 2048 values per tile
constexpr int kScanGroup = 32;                      // tiles per look-back group

struct ScanState {          // lives in the context workspace, reset to 0xFF.. per launch
    unsigned int *ticket;   // dynamic tile id (starts at 0xFFFFFFFF -> first tile is 0)
    unsigned long long *agg;   // [tiles]   tile aggregates (bit pattern of a double)
    unsigned long long *cpref; // [groups+1] group prefixes C_g
};

inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }
inline size_t scan_state_bytes(int64_t n) {
    int64_t t = scan_tiles(n);
    return 16 + 8 * (size_t)t + 8 * (size_t)(t / kScanGroup + 2);
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ unsigned long long canon_bits(double v) {
    if (v != v) return 0x7FF8000000000000ull;  // canonical NaN can never equal the sentinel
    return (unsigned long long)__double_as_longlong(v);
}
// int64 scans (offspring counts of residual resampling): values >= 0, never all-ones
__device__ __forceinline__ unsigned long long canon_bits(long long v) { return (unsigned long long)v; }
template <typename T> __device__ __forceinline__ T from_bits(unsigned long long b);
template <> __device__ __forceinline__ double from_bits<double>(unsigned long long b) {
    return __longlong_as_double((long long)b);
}
template <> __device__ __forceinline__ long long from_bits<long long>(unsigned long long b) {
    return (long long)b;
}
template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return b < a ? b : a; }
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a < b ? b : a; }
__device__ __forceinline__ void store_items(double *out, int64_t i0, const double (&o)[8]) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) st2(out + i0 + j, o[j], o[j + 1]);
}
__device__ __forceinline__ void store_items(long long *out, int64_t i0, const long long (&o)[8]) {
#pragma unroll
    for (int j = 0; j < 8; j += 2)
        *reinterpret_cast<longlong2 *>(out + i0 + j) = make_longlong2(o[j], o[j + 1]);
}

// inclusive warp scan (Kogge-Stone) followed by a running max: non-decreasing in lane
template <typename T>
__device__ __forceinline__ T warp_scan_monotone(T v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = v + o;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = tmax(v, o);
    }
    return v;
}

// LOAD: struct with  __device__ void operator()(int64_t i0, int64_t n, T (&v)[8]) const
// filling v[j] with the value at index i0 + j (0 beyond n).  T = double or long long.
template <typename T, typename LOAD>
__device__ __forceinline__ void scan_tiles_loop(const LOAD &load, int64_t n, T *out,
                                               ScanState st) {
    __shared__ T s_warp[kBlock / 32];
    __shared__ T s_pref[2];
    __shared__ unsigned int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int ntiles = (unsigned int)((n + kScanTile - 1) / kScanTile);

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(st.ticket, 1u) + 1u;
        __syncthreads();
        const unsigned int tile = s_tile;
        if (tile >= ntiles) break;
        const int64_t i0 = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;

        T r[kScanItems];
        load(i0, n, r);
#pragma unroll
        for (int j = 1; j < kScanItems; j++) r[j] = r[j - 1] + r[j];  // thread-local running sums

        // block scan of the thread totals
        T iw = warp_scan_monotone(r[kScanItems - 1], lane);
        if (lane == 31) s_warp[warp] = iw;
        __syncthreads();
        T woff = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++)
            if (w < warp) woff = woff + s_warp[w];
        T incl = woff + iw;                                  // I_k
        T up = __shfl_up_sync(0xffffffffu, iw, 1);
        T excl = (lane == 0) ? woff : (woff + up);           // E_k = I_{k-1}
        if (tid == kBlock - 1)                                    // publish the tile aggregate
            *reinterpret_cast<volatile unsigned long long *>(st.agg + tile) = canon_bits(incl);

        // look-back: warp 0 rebuilds this tile's exclusive prefix P_t and P_{t+1}
        if (warp == 0) {
            const unsigned int g = tile / kScanGroup, rr = tile % kScanGroup;
            // (1) nearest published group prefix C_k, k <= g: 32 candidates per probe round
            //     (lane l looks at C_{base-l}); C_0 = 0 is published by definition.
            unsigned int k = 0;
            T c = 0;
            for (int base = (int)g;; base -= 32) {
                const int idx = base - lane;
                unsigned long long b = kNotReady;
                if (idx >= 1) b = ld_volatile_u64(st.cpref + idx);
                else if (idx == 0) b = canon_bits((T)0);
                const unsigned int ready = __ballot_sync(0xffffffffu, b != kNotReady);
                if (ready) {
                    const int first = __ffs(ready) - 1;          // smallest lane = largest index
                    k = (unsigned int)(base - first);
                    c = from_bits<T>(__shfl_sync(0xffffffffu, b, first));
                    break;
                }
            }
            // (2) fold whole groups k .. g-1 (all their tiles precede ours): the loads of up to
            //     4 groups are issued together, the fold itself is sequential in a fixed order.
            for (; k < g; k += 4) {
                unsigned long long b4[4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                    b4[i] = (k + i < g) ? ld_volatile_u64(st.agg + (size_t)(k + i) * kScanGroup + lane) : 0ull;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (k + i < g) {
                        while (b4[i] == kNotReady)
                            b4[i] = ld_volatile_u64(st.agg + (size_t)(k + i) * kScanGroup + lane);
                        T sg = warp_scan_monotone(from_bits<T>(b4[i]), lane);
                        sg = __shfl_sync(0xffffffffu, sg, 31);
                        c = c + sg;
                        if (lane == 0)
                            *reinterpret_cast<volatile unsigned long long *>(st.cpref + k + i + 1) = canon_bits(c);
                    }
                }
            }
            T a = 0;
            if ((unsigned int)lane < rr) {
                unsigned long long b;
                do { b = ld_volatile_u64(st.agg + (size_t)g * kScanGroup + lane); } while (b == kNotReady);
                a = from_bits<T>(b);
            } else if ((unsigned int)lane == rr) {   // own aggregate: same association as `incl`
                T own = 0;                         // of thread kBlock-1, no global round trip
#pragma unroll
                for (int w = 0; w < kBlock / 32; w++) own = own + s_warp[w];
                a = from_bits<T>(canon_bits(own));
            }
            T ig = warp_scan_monotone(a, lane);
            T i_prev = __shfl_sync(0xffffffffu, ig, rr > 0 ? rr - 1 : 0);
            T i_this = __shfl_sync(0xffffffffu, ig, rr);
            T p_t = rr > 0 ? (c + i_prev) : c;
            T p_next = c + i_this;
            if (lane == 0) {
                s_pref[0] = p_t;
                s_pref[1] = p_next;
                if (rr == kScanGroup - 1)
                    *reinterpret_cast<volatile unsigned long long *>(st.cpref + g + 1) = canon_bits(p_next);
            }
        }
        __syncthreads();
        const T p_t = s_pref[0], p_next = s_pref[1];
        const T base = p_t + excl;
        const T cap = tmin(p_t + incl, p_next);
        T o[kScanItems];
#pragma unroll
        for (int j = 0; j < kScanItems; j++) o[j] = tmin(base + r[j], cap);
        if (i0 + kScanItems <= n) {
            store_items(out, i0, o);
        } else {
#pragma unroll
            for (int j = 0; j < kScanItems; j++)
                if (i0 + j < n) out[i0 + j] = o[j];
        }
        __syncthreads();  // s_tile / s_pref / s_warp are reused by the next tile
    }
}

}  // namespace smcb
