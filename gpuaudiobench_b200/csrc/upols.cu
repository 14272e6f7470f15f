// upols.cu — uniformly-partitioned overlap-save convolution for sm_100a (no cuFFT).
//
// Replaces the reference's Conv1D_accel pipeline (cuda/bench_conv1d_accel.cu:258-304: memset,
// T device-to-device memcpys, one whole-IR-length cufftExecR2C, ComplexMultiplyKernel :9-30,
// cufftExecC2R, ExtractRealPartKernel :32-47 — stateless, FFT size nextpow2(L+B-1) per buffer)
// with the streaming scheme of SURVEY.md App. E ("UPOLS engine"):
//
//   N = 2B, P = ceil(L/B);  H_p = RFFT_N([h[pB..pB+B) | 0]) / N        (setup, rfft_fwd_kernel)
//   X_m = RFFT_N([x_{m-1} | x_m])  -> ring slot (-m mod P)              (rfft_fwd_kernel)
//   Y_m[k] = sum_p H_p[k] * X_{m-p}[k]                                  (fdl_mac_kernel, HBM-bound)
//   y_m = IRFFT_N(Y_m)[B..2B)                                           (irfft_ols_kernel)
//
// Spectra hold B packed bins: bin 0 carries {DC, Nyquist} (both real), so a partition is exactly
// B*8 bytes — a whole number of 128 B lines for every supported B — and the MAC streams long
// contiguous runs.  The real transforms are done as one complex FFT of size M = B on the
// even/odd-packed signal (shared-memory Stockham autosort, radix 4 with one radix-2 pass when
// log2 M is odd) plus a twiddle post-/pre-pass.  Twiddles are computed in-kernel (sincospif, exact
// arguments) so the FFT phases carry no dependent table loads.
#include "upols.cuh"

#include <cstdlib>

#include "common.cuh"

namespace b200conv {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// e^{-2 pi i q / n} for integers 0 <= q, n a power of two: the argument of sincospif is exact, the
// result is within 1 ulp, and — unlike a table — it costs no dependent memory access (the FFT phases
// of the fused kernel are latency chains while HBM sits idle).
__device__ __forceinline__ float2 twiddle(int q, float inv_n) {
    float sn, cs;
    sincospif(-2.0f * static_cast<float>(q) * inv_n, &sn, &cs);
    return make_float2(cs, sn);
}

// ---------------------------------------------------------------------------------------------
// Complex FFT of size M on shared memory (Stockham autosort, out-of-place between a and b).
// Pass with sub-transform length Ns and radix R, butterfly j < M/R:
//   v[r] = in[j + r*M/R] * W_M^{r * (j mod Ns) * M/(Ns*R)};  DFT_R(v);
//   out[(j - j mod Ns)*R + (j mod Ns) + r*Ns] = v[r]
// INV uses conjugate twiddles and the +i rotation (un-normalised inverse).
// Every thread of the CTA must call this (it contains __syncthreads); `active` masks the work.
// Returns the buffer that holds the result.
// ---------------------------------------------------------------------------------------------
template <bool INV>
__device__ float2* fft_stockham(float2* a, float2* b, int M, int logM, int tx, int TX, bool active) {
    const float inv_m = 1.0f / static_cast<float>(M);
    int Ns = 1;
    if (logM & 1) {
        if (active) {
            const int half = M >> 1;
            for (int j = tx; j < half; j += TX) {
                float2 v0 = a[j], v1 = a[j + half];
                b[2 * j] = cadd(v0, v1);
                b[2 * j + 1] = csub(v0, v1);
            }
        }
        __syncthreads();
        float2* tmp = a; a = b; b = tmp;
        Ns = 2;
    }
    const int quarter = M >> 2;
    for (; Ns < M; Ns <<= 2) {
        if (active) {
            const int tstride = M / (4 * Ns);
            for (int j = tx; j < quarter; j += TX) {
                const int k = j & (Ns - 1);
                float2 v0 = a[j];
                float2 v1 = a[j + quarter];
                float2 v2 = a[j + 2 * quarter];
                float2 v3 = a[j + 3 * quarter];
                if (k != 0) {
                    float2 w1 = twiddle(k * tstride, inv_m);
                    float2 w2 = twiddle(2 * k * tstride, inv_m);
                    float2 w3 = twiddle(3 * k * tstride, inv_m);
                    if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
                    v1 = cmul(v1, w1);
                    v2 = cmul(v2, w2);
                    v3 = cmul(v3, w3);
                }
                float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), d = csub(v1, v3);
                float2 t3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // (+/-) i * d
                const int j0 = ((j - k) << 2) + k;
                b[j0] = cadd(t0, t2);
                b[j0 + Ns] = cadd(t1, t3);
                b[j0 + 2 * Ns] = csub(t0, t2);
                b[j0 + 3 * Ns] = csub(t1, t3);
            }
        }
        __syncthreads();
        float2* tmp = a; a = b; b = tmp;
    }
    return a;
}

// threads per window / windows per CTA for a 256-thread CTA
static inline int fft_threads_per_window(int M) {
    int tx = M / 4;
    if (tx < 32) tx = 32;
    if (tx > 256) tx = 256;
    return tx;
}

// ---------------------------------------------------------------------------------------------
// Forward: packed bins of RFFT_N([first | second]).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rfft_fwd_kernel(RfftParams p, int TX, int WPC) {
    extern __shared__ __align__(16) float2 fsm[];
    const int M = p.M, half = M >> 1;
    const int wl = threadIdx.x / TX, tx = threadIdx.x % TX;
    const int w = blockIdx.x * WPC + wl;
    const bool active = (w < p.count);
    float2* a = fsm + static_cast<size_t>(wl) * 2 * M;
    float2* b = a + M;

    if (active) {
        const float2* f2 = p.first ? reinterpret_cast<const float2*>(p.first + static_cast<size_t>(w) * p.first_stride) : nullptr;
        const float2* s2 = p.second ? reinterpret_cast<const float2*>(p.second + static_cast<size_t>(w) * p.second_stride) : nullptr;
        for (int n = tx; n < M; n += TX) {
            float2 v = make_float2(0.0f, 0.0f);
            if (n < half) {
                if (f2) v = f2[n];
            } else {
                if (s2) v = s2[n - half];
            }
            a[n] = v;
        }
    }
    __syncthreads();
    // prev_out may alias `first` (the engine's previous-buffer row): write it only after every lane of the
    // window has read its half — for M < 64 the reader of f2[k] and the writer of prev2[k] are different
    // lanes of one warp on divergent sides of the branch above.  The first FFT pass only reads a.
    if (active && p.prev_out) {
        float2* prev2 = reinterpret_cast<float2*>(p.prev_out + static_cast<size_t>(w) * M);
        for (int n = tx; n < half; n += TX) prev2[n] = a[n + half];
    }
    const float2* z = fft_stockham<false>(a, b, M, p.logM, tx, TX, active);
    if (!active) return;

    float2* out = p.out + static_cast<size_t>(w) * p.out_stride;
    const float sc = p.scale;
    for (int k = tx; k <= half; k += TX) {
        const float2 A = z[k];
        const float2 Bc = cconj(z[(M - k) & (M - 1)]);
        const float2 E = make_float2(0.5f * (A.x + Bc.x), 0.5f * (A.y + Bc.y));
        const float2 D = make_float2(0.5f * (A.x - Bc.x), 0.5f * (A.y - Bc.y));
        const float2 O = make_float2(D.y, -D.x);  // -i * D
        const float2 WO = cmul(twiddle(k, 0.5f / static_cast<float>(M)), O);
        const float2 Xk = cadd(E, WO);
        const float2 Xmk = make_float2(E.x - WO.x, -(E.y - WO.y));  // X[M-k] = conj(E - W O)
        if (k == 0) {
            if (p.unpacked) {
                out[0] = make_float2(sc * Xk.x, 0.0f);   // DC
                out[M] = make_float2(sc * Xmk.x, 0.0f);  // Nyquist
            } else {
                out[0] = make_float2(sc * Xk.x, sc * Xmk.x);  // {DC, Nyquist}
            }
        } else {
            out[k] = make_float2(sc * Xk.x, sc * Xk.y);
            if (k != half) out[M - k] = make_float2(sc * Xmk.x, sc * Xmk.y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// FDL-MAC: the HBM-streaming kernel.  Each thread owns two bins (one float4) of one track and
// walks the partitions of its split; H and the X ring are each read exactly once per block.
// Per bin the four real products are accumulated separately (A = sum Hr Xr, B = sum Hi Xi,
// C = sum Hr Xi, D = sum Hi Xr): re = A - B, im = C + D, and for packed bin 0 the pair (A, B) is
// directly {DC, Nyquist} — no divergent special case, same FMA count as a complex MAC.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mac2(float (&acc)[8], const float4& h, const float4& x) {
    acc[0] = fmaf(h.x, x.x, acc[0]);
    acc[1] = fmaf(h.y, x.y, acc[1]);
    acc[2] = fmaf(h.x, x.y, acc[2]);
    acc[3] = fmaf(h.y, x.x, acc[3]);
    acc[4] = fmaf(h.z, x.z, acc[4]);
    acc[5] = fmaf(h.w, x.w, acc[5]);
    acc[6] = fmaf(h.z, x.w, acc[6]);
    acc[7] = fmaf(h.w, x.z, acc[7]);
}

__global__ void __launch_bounds__(256, 8) fdl_mac_kernel(MacParams p, int U, int G) {
    __shared__ float red[256 * 8];
    const int t = blockIdx.z, s = blockIdx.y;
    int u, g;
    if (G == 1) {
        u = blockIdx.x * 256 + threadIdx.x;
        g = 0;
    } else {
        u = threadIdx.x % U;
        g = threadIdx.x / U;
    }
    const int P = p.P;
    const int p0 = static_cast<int>(static_cast<long long>(P) * s / p.S);
    const int p1 = static_cast<int>(static_cast<long long>(P) * (s + 1) / p.S);
    const float4* H4 = reinterpret_cast<const float4*>(p.H) + static_cast<size_t>(t) * P * U + u;
    const float4* X4 = reinterpret_cast<const float4*>(p.X) + static_cast<size_t>(t) * P * U + u;

    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;

    int pp = p0 + g;
    for (; pp + 3 * G < p1; pp += 4 * G) {
        float4 h[4], x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = pp + j * G;
            int sl = p.slot0 + q;
            if (sl >= P) sl -= P;
            h[j] = ldg_stream(H4 + static_cast<size_t>(q) * U);
            x[j] = ldg_stream(X4 + static_cast<size_t>(sl) * U);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) mac2(acc, h[j], x[j]);
    }
    for (; pp < p1; pp += G) {
        int sl = p.slot0 + pp;
        if (sl >= P) sl -= P;
        float4 h = ldg_stream(H4 + static_cast<size_t>(pp) * U);
        float4 x = ldg_stream(X4 + static_cast<size_t>(sl) * U);
        mac2(acc, h, x);
    }

    if (G > 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = acc[i];
        __syncthreads();
        if (g != 0) return;
        for (int gg = 1; gg < G; ++gg) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += red[(gg * U + u) * 8 + i];
        }
    }
    float4 y;
    if (u == 0) {
        y.x = acc[0];  // DC
        y.y = acc[1];  // Nyquist
    } else {
        y.x = acc[0] - acc[1];
        y.y = acc[2] + acc[3];
    }
    y.z = acc[4] - acc[5];
    y.w = acc[6] + acc[7];
    reinterpret_cast<float4*>(p.Ypart)[(static_cast<size_t>(s) * p.T + t) * U + u] = y;
}

// ---------------------------------------------------------------------------------------------
// Inverse + overlap-save: y = IRFFT_N(sum_s Ypart)[B..2B).  H carries the 1/N, so the inverse is
// un-normalised and Z'[k] = (Y[k] + conj Y[M-k]) + i conj(W^k) (Y[k] - conj Y[M-k]).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) irfft_ols_kernel(IrfftParams p, int TX, int WPC) {
    extern __shared__ __align__(16) float2 fsm[];
    const int M = p.M, half = M >> 1;
    const int wl = threadIdx.x / TX, tx = threadIdx.x % TX;
    const int t = blockIdx.x * WPC + wl;
    const bool active = (t < p.T);
    float2* a = fsm + static_cast<size_t>(wl) * 2 * M;
    float2* b = a + M;

    if (active) {
        const size_t split_stride = static_cast<size_t>(p.T) * M;
        const float2* Y = p.Ypart + static_cast<size_t>(t) * M;
        for (int k = tx; k <= half; k += TX) {
            const int mk = (M - k) & (M - 1);
            float2 yk = Y[k], ym = Y[mk];
            for (int s = 1; s < p.S; ++s) {
                yk = cadd(yk, Y[s * split_stride + k]);
                ym = cadd(ym, Y[s * split_stride + mk]);
            }
            float2 A, Bm;
            if (k == 0) {
                A = make_float2(yk.x, 0.0f);   // DC
                Bm = make_float2(yk.y, 0.0f);  // Nyquist = Y[M]
            } else {
                A = yk;
                Bm = ym;
            }
            const float2 Bc = cconj(Bm);
            const float2 E = cadd(A, Bc);
            const float2 D = csub(A, Bc);
            const float2 O = cmul(cconj(twiddle(k, 0.5f / static_cast<float>(M))), D);
            a[k] = make_float2(E.x - O.y, E.y + O.x);
            if (k != 0 && k != half) a[M - k] = make_float2(E.x + O.y, O.x - E.y);
        }
    }
    __syncthreads();
    const float2* z = fft_stockham<true>(a, b, M, p.logM, tx, TX, active);
    if (!active) return;

    if (!p.sample_major) {
        float2* out2 = reinterpret_cast<float2*>(p.out + static_cast<size_t>(t) * M);
        for (int n = tx; n < half; n += TX) out2[n] = z[half + n];
    } else {
        float* col = p.out + p.toff + t;
        for (int n = tx; n < half; n += TX) {
            const float2 v = z[half + n];
            col[static_cast<size_t>(2 * n) * p.Tg] = v.x;
            col[static_cast<size_t>(2 * n + 1) * p.Tg] = v.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fused block kernel (M = B <= 2048): forward FFT -> FDL-MAC -> inverse FFT + overlap-save in ONE
// launch; grid (S * KT, T), 256 threads.  KT = max(1, M/512) bin tiles of 512 bins: for M > 512 a track's
// spectrum is wider than the 256 float4 a CTA's threads cover, so the MAC of a (split, tile) pair is a CTA of
// its own (the MAC is separable by bins; round 1 ran the 3-kernel path here at 0.63-0.81 of the HBM peak on the
// step).  Tile 0 of split 0 runs the forward transform and publishes X_m to the ring; the other tiles of split
// 0 stream their older partitions first and take partition 0 last, from the ring, once tile 0 has raised the
// track's "spectrum ready" word (the first version let every tile transform the window itself: 8 redundant
// 4096-point FFTs and 8 PCIe reads of the input per track, 192 -> 285 us at B = 4096).  The newest spectrum X_m only enters the sum through
// partition 0, which belongs to split 0, so only that CTA transforms the input (and publishes
// X_m to the ring for later blocks); the other splits stream older ring slots straight away.
// Each CTA leaves its partial spectrum in Ypart and takes a ticket on the track's counter; the
// last one adds the S partials in split order, runs the inverse transform and writes the output.
// ---------------------------------------------------------------------------------------------
// Occupancy variants <UNROLL, MINCTAS> with UNROLL * MINCTAS = 16, i.e. the same bytes in flight per SM at
// full residency (256 threads x UNROLL x 2 loads x 16 B x MINCTAS = 131 KB).  Measured (C3: 1024 CTAs,
// C4 shard: 512 CTAs; profiles/experiments/upols_variant_bench.py): a grid that needs ~1.7 waves beats one
// that is exactly resident, because every CTA starts and ends with a transform phase that moves no HBM
// bytes and in a single wave those phases line up across the whole GPU — C3 <2,8> 171.1 us (1 wave)
// vs <4,4> 161.9 us (1.73 waves); C4 <2,8> 127.9 us vs <8,2> 125.1 us vs <4,4> with the partition range
// split in 2 / 3 (1024 / 1536 CTAs) 124.0 / 122.5 us.  The product uses <4,4> and the planner
// (engine.cu plan_upols) sizes the split for about two waves.
#ifndef B200CONV_FUSED_PREFETCH
#define B200CONV_FUSED_PREFETCH 8
#endif
constexpr int kPrefetchParts = B200CONV_FUSED_PREFETCH;  // partitions pulled into L2 under the forward FFT

template <int kFusedUnroll, int kMinCtas, bool kStrip, bool kBusTree = false>
__global__ void __launch_bounds__(256, kMinCtas) upols_fused_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(16) float2 fsm[];  // [2][M] FFT ping-pong | red[256*8]
    __shared__ int s_last;
    const int M = p.M, half = M >> 1, U = M >> 1;  // U: float4 (bin pairs) per partition row
    const int KT = p.KT;                           // bin tiles (1 for M <= 512)
    const int G = (U >= 256) ? 1 : 256 / U;        // partition lanes per bin pair
    const int s = blockIdx.x / KT, kt = blockIdx.x - s * KT, t = blockIdx.y, tid = threadIdx.x;
    float2* a = fsm;
    float2* b = fsm + M;
    float* red = reinterpret_cast<float*>(fsm + 2 * M);
    const float2* xsm = nullptr;  // packed X_m in shared memory (split 0 only)
    pdl_launch_dependents();      // the bus kernel may be scheduled as our CTAs retire

    // Keep HBM busy while this CTA runs its forward FFT: pull the first kPrefetchParts partitions of
    // its H rows and ring slots into L2 (no registers, no waiting); the MAC loop then finds them there.
    {
        const int Pp = p.P;
        const int pf0 = static_cast<int>(static_cast<long long>(Pp) * s / p.S);
        const int pf1 = static_cast<int>(static_cast<long long>(Pp) * (s + 1) / p.S);
        const int lines_per_row = (min(U, 256) * 16) >> 7;  // 128-byte lines of this CTA's bin tile in a partition row
        const int tile_off = kt * 256 * 16;                  // byte offset of the tile inside a row
        const int total = min(kPrefetchParts, pf1 - pf0) * lines_per_row;
        for (int i = tid; i < total; i += 256) {
            const int q = pf0 + i / lines_per_row, ln = i - (i / lines_per_row) * lines_per_row;
            int sl = p.slot0 + q;
            if (sl >= Pp) sl -= Pp;
            const char* hrow = reinterpret_cast<const char*>(p.H + (static_cast<size_t>(t) * Pp + q) * M) + tile_off + ln * 128;
            const char* xrow = reinterpret_cast<const char*>(p.X + (static_cast<size_t>(t) * Pp + sl) * M) + tile_off + ln * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(hrow));
            if (q != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow));  // slot0 itself is written below
        }
    }

    if (s == 0 && kt == 0) {
        const float2* in2 = reinterpret_cast<const float2*>(p.d_in + static_cast<size_t>(t) * M);
        const float2* prev2 = reinterpret_cast<const float2*>(p.prev + static_cast<size_t>(t) * M);
        float2* prevw2 = reinterpret_cast<float2*>(p.prev_w + static_cast<size_t>(t) * M);  // the OTHER half of the ping-pong:
        const bool publish = true;
        for (int n = tid; n < half; n += 256) {
            const float2 pv = prev2[n], cv = in2[n];
            a[n] = pv;          // window = [previous buffer | current buffer], even/odd packed
            a[n + half] = cv;
            if (p.commit && publish) prevw2[n] = cv;
        }
        __syncthreads();
        float2* z = fft_stockham<false>(a, b, M, p.logM, tid, 256, true);
        float2* xo = (z == a) ? b : a;
        float2* ring = p.X + (static_cast<size_t>(t) * p.P + p.slot0) * M;
        for (int k = tid; k <= half; k += 256) {
            const float2 A = z[k];
            const float2 Bc = cconj(z[(M - k) & (M - 1)]);
            const float2 E = make_float2(0.5f * (A.x + Bc.x), 0.5f * (A.y + Bc.y));
            const float2 D = make_float2(0.5f * (A.x - Bc.x), 0.5f * (A.y - Bc.y));
            const float2 WO = cmul(twiddle(k, 0.5f / static_cast<float>(M)), make_float2(D.y, -D.x));
            const float2 Xk = cadd(E, WO);
            const float2 Xmk = make_float2(E.x - WO.x, -(E.y - WO.y));
            if (k == 0) {
                xo[0] = make_float2(Xk.x, Xmk.x);
                if (publish) ring[0] = xo[0];
            } else {
                xo[k] = Xk;
                if (publish) ring[k] = Xk;
                if (k != half) {
                    xo[M - k] = Xmk;
                    if (publish) ring[M - k] = Xmk;
                }
            }
        }
        __syncthreads();
        xsm = xo;
        if (KT > 1) {  // the other bin tiles of this track take partition 0 from the ring: tell them it is there
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(&p.xready[t], p.seq);
        }
    }

    // ---- FDL-MAC over this split's partitions (same loop as fdl_mac_kernel) ----
    const int u = (U >= 256) ? kt * 256 + tid : tid % U;  // this thread's bin pair inside a partition row
    const int g = (U >= 256) ? 0 : tid / U;
    const int P = p.P;
    int p0 = static_cast<int>(static_cast<long long>(P) * s / p.S);
    const int p1 = static_cast<int>(static_cast<long long>(P) * (s + 1) / p.S);
    const float4* H4 = reinterpret_cast<const float4*>(p.H) + static_cast<size_t>(t) * P * U + u;
    const float4* X4 = reinterpret_cast<const float4*>(p.X) + static_cast<size_t>(t) * P * U + u;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    if (s == 0) {
        // partition 0 against the spectrum still in shared memory (tile 0) / from the ring after the older ones (other tiles)
        if (kt == 0 && g == 0) mac2(acc, ldg_stream(H4), reinterpret_cast<const float4*>(xsm)[u]);
        p0 = 1;
    }
    int pp = p0 + g;
    for (; pp + (kFusedUnroll - 1) * G < p1; pp += kFusedUnroll * G) {
        float4 h[kFusedUnroll], x[kFusedUnroll];
#pragma unroll
        for (int j = 0; j < kFusedUnroll; ++j) {
            const int q = pp + j * G;
            int sl = p.slot0 + q;
            if (sl >= P) sl -= P;
            h[j] = ldg_stream(H4 + static_cast<size_t>(q) * U);
            x[j] = ldg_stream(X4 + static_cast<size_t>(sl) * U);
        }
#pragma unroll
        for (int j = 0; j < kFusedUnroll; ++j) mac2(acc, h[j], x[j]);
    }
    for (; pp < p1; pp += G) {
        int sl = p.slot0 + pp;
        if (sl >= P) sl -= P;
        mac2(acc, ldg_stream(H4 + static_cast<size_t>(pp) * U), ldg_stream(X4 + static_cast<size_t>(sl) * U));
    }
    if (s == 0 && kt != 0) {  // (KT > 1 implies G == 1) partition 0 now: wait for tile 0's transform of this block
        if (tid == 0) {
            unsigned spins = 0;
            while (atomicAdd(&p.xready[t], 0u) != p.seq && ++spins < (1u << 24)) {
            }
            __threadfence();
        }
        __syncthreads();
        const float4 h0 = ldg_stream(H4);
        const float4 x0 = __ldcg(X4 + static_cast<size_t>(p.slot0) * U);  // written by another CTA of this launch: L2, not L1
        mac2(acc, h0, x0);
    }
    if (G > 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[tid * 8 + i] = acc[i];
        __syncthreads();
        if (g == 0) {
            for (int gg = 1; gg < G; ++gg) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += red[(gg * U + u) * 8 + i];
            }
        }
    }
    float4 y;
    if (u == 0) {
        y.x = acc[0];  // DC
        y.y = acc[1];  // Nyquist
    } else {
        y.x = acc[0] - acc[1];
        y.y = acc[2] + acc[3];
    }
    y.z = acc[4] - acc[5];
    y.w = acc[6] + acc[7];

    __syncthreads();  // everyone is done with xsm / the FFT buffers
    float2* ysm = b;  // summed spectrum goes to b, the inverse pre-pass writes a
    if (p.S * KT == 1) {
        if (g == 0) reinterpret_cast<float4*>(ysm)[u] = y;
    } else {
        if (g == 0) reinterpret_cast<float4*>(p.Ypart)[(static_cast<size_t>(s) * p.T + t) * U + u] = y;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned ticket = atomicAdd(&p.counters[t], 1u);
            s_last = (ticket == static_cast<unsigned>(p.S * KT) - 1u);  // every (split, bin tile) CTA of the track
            if (s_last) p.counters[t] = 0;  // re-armed for the next block
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        // the whole spectrum of the track (all bin tiles), splits added in split order
        for (int uu = (U >= 256) ? tid : u; uu < U; uu += 256) {
            if (g == 0) {
                float4 sum = __ldcg(reinterpret_cast<const float4*>(p.Ypart) + static_cast<size_t>(t) * U + uu);
                for (int ss = 1; ss < p.S; ++ss) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(p.Ypart) + (static_cast<size_t>(ss) * p.T + t) * U + uu);
                    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                }
                reinterpret_cast<float4*>(ysm)[uu] = sum;
            }
        }
    }
    __syncthreads();

    // ---- inverse real-FFT pre-pass, inverse FFT, overlap-save output ----
    for (int k = tid; k <= half; k += 256) {
        const float2 yk = ysm[k], ym = ysm[(M - k) & (M - 1)];
        float2 A, Bm;
        if (k == 0) {
            A = make_float2(yk.x, 0.0f);
            Bm = make_float2(yk.y, 0.0f);
        } else {
            A = yk;
            Bm = ym;
        }
        const float2 Bc = cconj(Bm);
        const float2 E = cadd(A, Bc);
        const float2 O = cmul(cconj(twiddle(k, 0.5f / static_cast<float>(M))), csub(A, Bc));
        a[k] = make_float2(E.x - O.y, E.y + O.x);
        if (k != 0 && k != half) a[M - k] = make_float2(E.x + O.y, O.x - E.y);
    }
    __syncthreads();
    const float2* z = fft_stockham<true>(a, b, M, p.logM, tid, 256, true);
    if (kStrip) {
        // channel strip on the track's B output samples while they are still in shared memory: one
        // thread walks the dependent chain (~12 cycles per sample) while the SM's other CTAs keep streaming
        __syncthreads();
        if (tid == 0) strip_track_in_smem(p.strip, t, const_cast<float*>(reinterpret_cast<const float*>(z + half)), M);
        __syncthreads();
    }
    if (kBusTree && p.bus.mix) {  // measurement option (B200CONV_BUS_TREE=1): a separate instantiation, the product kernel has none of this
        // stereo bus first (its tickets are the critical path after the last track): this track's row goes to the
        // tree's scratch; the last track of a group sums the group, groups are folded in order into the running bus,
        // and on a multi-GPU job the last one exchanges the bus over NVLink (bus_tree.cuh) — no further launch
        float2* yb = reinterpret_cast<float2*>(p.bus.ybus + static_cast<size_t>(t) * M);
        for (int n = tid; n < half; n += 256) yb[n] = z[half + n];
        bus_tree_arrive(p.bus, t, 0, tid, 256, 0, &s_last);
    }
    for (int copy = 0; copy < 2; ++copy) {
        float* dst = copy ? p.out2 : p.out;
        if (!dst) continue;
        if (!p.sample_major) {
            float2* o2 = reinterpret_cast<float2*>(dst + static_cast<size_t>(t) * M);
            for (int n = tid; n < half; n += 256) o2[n] = z[half + n];
        } else {
            float* col = dst + p.toff + t;
            for (int n = tid; n < half; n += 256) {
                const float2 v = z[half + n];
                col[static_cast<size_t>(2 * n) * p.Tg] = v.x;
                col[static_cast<size_t>(2 * n + 1) * p.Tg] = v.y;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
static cudaError_t ensure_fft_smem(const void* fn, size_t smem) { return ensure_dyn_smem(fn, smem); }

cudaError_t launch_rfft_fwd(const RfftParams& p, cudaStream_t st) {
    const int TX = fft_threads_per_window(p.M);
    const int WPC = 256 / TX;
    const size_t smem = static_cast<size_t>(WPC) * 2 * p.M * sizeof(float2);
    cudaError_t e = ensure_fft_smem(reinterpret_cast<const void*>(rfft_fwd_kernel), smem);
    if (e != cudaSuccess) return e;
    rfft_fwd_kernel<<<(p.count + WPC - 1) / WPC, 256, smem, st>>>(p, TX, WPC);
    return cudaGetLastError();
}

cudaError_t launch_irfft_ols(const IrfftParams& p, cudaStream_t st) {
    const int TX = fft_threads_per_window(p.M);
    const int WPC = 256 / TX;
    const size_t smem = static_cast<size_t>(WPC) * 2 * p.M * sizeof(float2);
    cudaError_t e = ensure_fft_smem(reinterpret_cast<const void*>(irfft_ols_kernel), smem);
    if (e != cudaSuccess) return e;
    irfft_ols_kernel<<<(p.T + WPC - 1) / WPC, 256, smem, st>>>(p, TX, WPC);
    return cudaGetLastError();
}

cudaError_t launch_fdl_mac(const MacParams& p, cudaStream_t st) {
    const int U = p.M / 2;
    int G = 1, KT = 1;
    if (U >= 256)
        KT = U / 256;
    else
        G = 256 / U;
    dim3 grid(KT, p.S, p.T);
    fdl_mac_kernel<<<grid, 256, 0, st>>>(p, U, G);
    return cudaGetLastError();
}

// CTAs per SM of the fused kernel: <4,4> measured best at every grid size tried (C3 with 1024 and 2048
// CTAs, C4 shard with 512, 1024 and 1536); the other two instantiations stay selectable for experiments.
int upols_fused_occupancy() {
    static const int forced = [] {
        const char* v = std::getenv("B200CONV_UPOLS_OCC");
        return v ? std::atoi(v) : 0;
    }();
    return (forced == 8 || forced == 2) ? forced : 4;
}

template <typename K>
static cudaError_t launch_fused_variant(K kernel, const FusedParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kernel), smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_upols_fused(const FusedParams& p, cudaStream_t st) {
    dim3 grid(p.S * p.KT, p.T);
    const size_t smem = static_cast<size_t>(2) * p.M * sizeof(float2) + 256 * 8 * sizeof(float);
    if (p.bus.mix) {  // in-kernel bus tree requested (B200CONV_BUS_TREE=1)
        if (p.strip.ops) return launch_fused_variant(upols_fused_kernel<4, 4, true, true>, p, grid, smem, st);
        return launch_fused_variant(upols_fused_kernel<4, 4, false, true>, p, grid, smem, st);
    }
    // strip in the epilogue: one more instantiation of the product configuration
    if (p.strip.ops) return launch_fused_variant(upols_fused_kernel<4, 4, true>, p, grid, smem, st);
    switch (upols_fused_occupancy()) {
        case 8: return launch_fused_variant(upols_fused_kernel<2, 8, false>, p, grid, smem, st);
        case 4: return launch_fused_variant(upols_fused_kernel<4, 4, false>, p, grid, smem, st);
        default: return launch_fused_variant(upols_fused_kernel<8, 2, false>, p, grid, smem, st);
    }
}

}  // namespace b200conv
