// jbugs_dense.cuh -- dense linear predictor  eta = X * beta  across chains as FP64 tensor-core GEMMs (DMMA),
// fused with the Bernoulli / binomial-logit log density and its reverse pass.
//
// Replaces, for a statement like  `y[i] ~ dbern(p[i]); logit(p[i]) <- inprod(X[i, 1:P], beta[1:P])`,
// one `inprod` node (JuliaBUGS/src/BUGSPrimitives/functions.jl:33-35) + one `logpdf` node per observation and per
// chain, and the AD backend's reverse sweep through them.
//
//   forward  : E[N x C]  = X[N x P] * B[P x C]          B = the beta rows of theta (chain-fast: row j = beta_j over chains)
//   epilogue : r = y - n * logistic(E),  logp_c += y*log p + (n-y)*log(1-p)      (fused into the forward GEMM; E is
//              never stored, R[N x C] is)
//   reverse  : G[P x C]  = X^T[P x N] * R[N x C]         split-K over the observations, partials added in a fixed order
//
// Both products run on the FP64 tensor cores with `mma.sync.aligned.m8n8k4.row.col.f64` (the sm_100a DMMA path;
// tcgen05 has no FP64 kind).  CTA tile 128 x 128 x 16, 16 warps (4 x 4), warp tile 32 x 32 = 4 x 4 MMA tiles (four
// warps per scheduler hide the DMMA issue cadence), three cp.async stages.  Shared-memory rows are padded to a stride
// == 4 (mod 16) doubles so that the 64-bit fragment loads of a half-warp hit 16 distinct bank pairs.  Interior tiles
// use unpredicated 16-byte cp.async with per-thread pointers that are only incremented in the k loop; edge tiles and
// unaligned operands take a predicated 8-byte zero-filling path.
#pragma once

#include "jbugs_kernels.cuh"

namespace jbugs {

#ifndef JB_DENSE_BWD_BK
#define JB_DENSE_BWD_BK 32  // slab depth of the reverse GEMM (16: round 1; 32 halves the barriers per k)
#endif
constexpr int DG_BN = 128, DG_STAGES = 3;
constexpr int DG_LDB = DG_BN + 4;   // Bs[k][n]
constexpr int DG_TM = 4, DG_TN = 4;  // MMA tiles per warp (32 x 32 warp tile)

// CTA shape: BM x 128 x 16 with BM / 32 x 4 warps.  The reverse GEMM uses 128 x 128 (16 warps, one CTA per SM: 86 % of
// the DMMA pipe); the forward GEMM uses 64 x 128 (8 warps, TWO CTAs per SM) so that the FP64 epilogue of one CTA --
// ~15 % of a tile's time, K = P is short -- runs under the DMMA main loop of the other.
template <int BM_, int BK_>
struct DgCfg {
    static constexpr int BM = BM_;
    static constexpr int BK = BK_;  // depth of one shared-memory slab
    static constexpr int THREADS = (BM_ / 32) * 4 * 32;
    static constexpr int WM = BM_ / 32;
    static constexpr int LDA_M = BM_ + 4;   // As[k][m]   (A is M-major in memory: forward GEMM)
    static constexpr int LDA_K = BK_ + 4; // As[m][k]   (A is K-major in memory: reverse GEMM)
    static constexpr int A_DOUBLES = (BK_ * LDA_M > BM_ * LDA_K) ? BK_ * LDA_M : BM_ * LDA_K;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + BK_ * DG_LDB;
    static constexpr size_t SMEM_BYTES = (size_t)DG_STAGES * STAGE_DOUBLES * sizeof(double);
    static constexpr int A_CH = BM_ * BK_ / 2 / THREADS;    // 16-byte chunks of A per thread and slab
    static constexpr int B_CH = BK_ * DG_BN / 2 / THREADS;  // 16-byte chunks of B per thread and slab
};
using DgFwd = DgCfg<64, 16>;
using DgBwd = DgCfg<128, JB_DENSE_BWD_BK>;
constexpr int DG_BM = DgBwd::BM;          // row tile of the reverse GEMM (over P)
constexpr int DG_FWD_BM = DgFwd::BM;      // row tile of the forward GEMM (over observations)

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

struct DenseOperands {
    const double* A;  // X
    long long lda;    // number of observations (leading dimension of the column-major X)
    const double* B;  // forward: theta + beta_base*ld (row j at j*ldb); reverse: R
    long long ldb;    // elements between consecutive k rows of B
    long long M, N, K;
    int vec_ok;       // operands are 16-byte aligned with even leading dimensions
};

// ---- edge path: one BK-slab, 8-byte elements, bounds-checked, zero fill ---------------------------------------------
// A_MMAJOR: A(m,k) = A[m + lda*k]; else A(m,k) = A[k + lda*m]
template <bool A_MMAJOR, class Cfg>
__device__ __forceinline__ void dg_load_stage_edge(const DenseOperands& op, double* stage, long long m0, long long n0,
                                                   long long k0, long long k_end, int tid) {
    double* As = stage;
    double* Bs = stage + Cfg::A_DOUBLES;
    for (int q = tid; q < Cfg::BK * Cfg::BM; q += Cfg::THREADS) {
        const int kk = A_MMAJOR ? q / Cfg::BM : q % Cfg::BK;
        const int m = A_MMAJOR ? q % Cfg::BM : q / Cfg::BK;
        double* dst = A_MMAJOR ? As + kk * Cfg::LDA_M + m : As + m * Cfg::LDA_K + kk;
        const bool ok = (m0 + m < op.M) && (k0 + kk < k_end);
        if (ok) cp_async8(dst, A_MMAJOR ? op.A + (m0 + m) + op.lda * (k0 + kk) : op.A + (k0 + kk) + op.lda * (m0 + m));
        else *dst = 0.0;
    }
    for (int q = tid; q < Cfg::BK * DG_BN; q += Cfg::THREADS) {
        const int kk = q / DG_BN, n = q % DG_BN;
        const bool ok = (n0 + n < op.N) && (k0 + kk < k_end);
        if (ok) cp_async8(Bs + kk * DG_LDB + n, op.B + (k0 + kk) * op.ldb + (n0 + n));
        else Bs[kk * DG_LDB + n] = 0.0;
    }
}

// ---- interior path: per-thread source pointers / shared offsets, two 16-byte chunks of A and two of B per slab ----------
template <class Cfg>
struct DgFastPtrs {
    const double* a[Cfg::A_CH];
    const double* b[Cfg::B_CH];
    int sa[Cfg::A_CH], sb[Cfg::B_CH];  // offsets (doubles) inside a stage
    long long a_step, b_step;
};

template <bool A_MMAJOR, class Cfg>
__device__ __forceinline__ DgFastPtrs<Cfg> dg_fast_init(const DenseOperands& op, long long m0, long long n0, long long k0, int tid) {
    DgFastPtrs<Cfg> f;
#pragma unroll
    for (int h = 0; h < Cfg::A_CH; ++h) {
        const int q = tid + h * Cfg::THREADS;  // chunks of 2 doubles
        if (A_MMAJOR) {
            const int kk = q / (Cfg::BM / 2), m = (q % (Cfg::BM / 2)) * 2;
            f.a[h] = op.A + (m0 + m) + op.lda * (k0 + kk);
            f.sa[h] = kk * Cfg::LDA_M + m;
        } else {
            const int m = q / (Cfg::BK / 2), kk = (q % (Cfg::BK / 2)) * 2;
            f.a[h] = op.A + (k0 + kk) + op.lda * (m0 + m);
            f.sa[h] = m * Cfg::LDA_K + kk;
        }
    }
#pragma unroll
    for (int h = 0; h < Cfg::B_CH; ++h) {
        const int q = tid + h * Cfg::THREADS;
        const int kb = q / (DG_BN / 2), n = (q % (DG_BN / 2)) * 2;
        f.b[h] = op.B + (k0 + kb) * op.ldb + (n0 + n);
        f.sb[h] = Cfg::A_DOUBLES + kb * DG_LDB + n;
    }
    f.a_step = A_MMAJOR ? op.lda * Cfg::BK : Cfg::BK;
    f.b_step = op.ldb * Cfg::BK;
    return f;
}

template <class Cfg>
__device__ __forceinline__ void dg_fast_load(DgFastPtrs<Cfg>& f, double* stage) {
#pragma unroll
    for (int h = 0; h < Cfg::A_CH; ++h) {
        cp_async16(stage + f.sa[h], f.a[h]);
        f.a[h] += f.a_step;
    }
#pragma unroll
    for (int h = 0; h < Cfg::B_CH; ++h) {
        cp_async16(stage + f.sb[h], f.b[h]);
        f.b[h] += f.b_step;
    }
}

// acc[tm][tn][2] += A_tile * B_tile over k in [k_begin, k_end)
template <bool A_MMAJOR, class Cfg>
__device__ __forceinline__ void dg_mainloop(const DenseOperands& op, double* smem, long long m0, long long n0,
                                            long long k_begin, long long k_end, double (&acc)[DG_TM][DG_TN][2]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // (BM / 32) x 4 warps
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index
    const long long nk = (k_end - k_begin + Cfg::BK - 1) / Cfg::BK;
    // interior tile with aligned operands: every slab but a ragged last one goes through the fast loader
    const bool interior = op.vec_ok && (m0 + Cfg::BM <= op.M) && (n0 + DG_BN <= op.N) && (((k_begin & 1) == 0) || A_MMAJOR);
    const long long nk_fast = interior ? (k_end - k_begin) / Cfg::BK : 0;
    DgFastPtrs<Cfg> fp = dg_fast_init<A_MMAJOR, Cfg>(op, m0, n0, k_begin, tid);
    auto load = [&](long long slab) {
        double* stage = smem + (slab % DG_STAGES) * Cfg::STAGE_DOUBLES;
        if (slab < nk_fast) dg_fast_load<Cfg>(fp, stage);
        else dg_load_stage_edge<A_MMAJOR, Cfg>(op, stage, m0, n0, k_begin + slab * Cfg::BK, k_end, tid);
    };
    for (int s = 0; s < DG_STAGES - 1; ++s) {
        if (s < nk) load(s);
        cp_async_commit();
    }
    for (long long kt = 0; kt < nk; ++kt) {
        cp_async_wait_1();
        __syncthreads();
        const long long pf = kt + DG_STAGES - 1;  // refill the stage consumed at iteration kt - 1
        if (pf < nk) load(pf);
        cp_async_commit();
        const double* As = smem + (kt % DG_STAGES) * Cfg::STAGE_DOUBLES;
        const double* Bs = As + Cfg::A_DOUBLES;
#pragma unroll
        for (int ks = 0; ks < Cfg::BK / 4; ++ks) {
            double a[DG_TM], b[DG_TN];
#pragma unroll
            for (int tm = 0; tm < DG_TM; ++tm) {
                const int m = wm * 32 + tm * 8 + fr, k = ks * 4 + fk;
                a[tm] = A_MMAJOR ? As[k * Cfg::LDA_M + m] : As[m * Cfg::LDA_K + k];
            }
#pragma unroll
            for (int tn = 0; tn < DG_TN; ++tn) b[tn] = Bs[(ks * 4 + fk) * DG_LDB + wn * 32 + tn * 8 + fr];
#pragma unroll
            for (int tm = 0; tm < DG_TM; ++tm)
#pragma unroll
                for (int tn = 0; tn < DG_TN; ++tn) dmma(acc[tm][tn], a[tm], b[tn]);
        }
    }
    cp_async_wait_all();
    __syncthreads();
}

// ---- forward GEMM + Bernoulli / binomial-logit epilogue --------------------------------------------------------
// grid: (col blocks of 128 chains, row groups); every CTA walks the row tiles rg, rg + gridDim.y, ... of its column block
// and keeps the per-chain log density in registers; partial[rg][c] is written at the end.
struct DenseFwdParams {
    DenseOperands op;         // A = X (M-major), B = beta rows of theta
    const double* y;          // [M]
    const double* n;          // [M] or null (Bernoulli: n = 1)
    double* R;                // [M x ldr] residuals y - n p
    long long ldr;
    double* partial;          // [gridDim.y][ldg]
    long long ldg;
    int* flags;
};

__global__ void __launch_bounds__(DgFwd::THREADS, 2)
dense_fwd_kernel(const __grid_constant__ DenseFwdParams P) {
    using Cfg = DgFwd;
    extern __shared__ double smem[];
    __shared__ double red[Cfg::WM][DG_BN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3, fr = lane >> 2, fk = lane & 3;
    const long long n0 = (long long)blockIdx.x * DG_BN;
    const long long m_tiles = (P.op.M + Cfg::BM - 1) / Cfg::BM;
    double lp[DG_TN][2];  // per-thread log-density sums for its 4 x 2 chain columns
#pragma unroll
    for (int tn = 0; tn < DG_TN; ++tn) lp[tn][0] = lp[tn][1] = 0.0;
    for (long long mt = blockIdx.y; mt < m_tiles; mt += gridDim.y) {
        const long long m0 = mt * Cfg::BM;
        double acc[DG_TM][DG_TN][2];
#pragma unroll
        for (int tm = 0; tm < DG_TM; ++tm)
#pragma unroll
            for (int tn = 0; tn < DG_TN; ++tn) acc[tm][tn][0] = acc[tm][tn][1] = 0.0;
        dg_mainloop<true, Cfg>(P.op, smem, m0, n0, 0, P.op.K, acc);
        // epilogue: C fragment element (row = fr, col = 2*fk + {0,1}) of MMA tile (tm, tn)
#pragma unroll
        for (int tm = 0; tm < DG_TM; ++tm) {
            const long long i = m0 + wm * 32 + tm * 8 + fr;
            const bool row_ok = i < P.op.M;
            const double yv = row_ok ? __ldg(P.y + i) : 0.0;
            const double nv = row_ok ? (P.n ? __ldg(P.n + i) : 1.0) : 0.0;
#pragma unroll
            for (int tn = 0; tn < DG_TN; ++tn) {
                double r2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double eta = acc[tm][tn][e];
                    // y log p + (n - y) log(1 - p) and y - n p, branch-free (jbugs_device.cuh: binomial_logit)
                    double l = 0.0, dl;
                    int sat = 0;
                    binomial_logit(eta, yv, nv, l, dl, sat);
                    if (sat && logit_tail_is_inf(eta, yv, nv)) l = -dev_inf();  // saturated tails (rare)
                    lp[tn][e] += row_ok ? l : 0.0;
                    r2[e] = dl;
                }
                const long long c = n0 + wn * 32 + tn * 8 + 2 * fk;
                if (row_ok && c < P.op.N) {
                    if (c + 1 < P.op.N) *reinterpret_cast<double2*>(P.R + i * P.ldr + c) = make_double2(r2[0], r2[1]);
                    else P.R[i * P.ldr + c] = r2[0];
                }
            }
        }
    }
    // fixed-order reduction of lp over the 8 fragment rows (shuffles), then over the four M-warps (shared memory)
#pragma unroll
    for (int tn = 0; tn < DG_TN; ++tn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double v = lp[tn][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (fr == 0) red[wm][wn * 32 + tn * 8 + 2 * fk + e] = v;
        }
    __syncthreads();
    if (tid < DG_BN) {
        const long long c = n0 + tid;
        if (c < P.op.N) {
            double v = red[0][tid];
#pragma unroll
            for (int w = 1; w < Cfg::WM; ++w) v += red[w][tid];  // fixed order over the M-warps
            P.partial[(long long)blockIdx.y * P.ldg + c] = v;
        }
    }
}

// ---- reverse GEMM, split-K -----------------------------------------------------------------------------------------
// grid: (col blocks, row blocks of P, k splits); Gp[ks][p][c]
struct DenseBwdParams {
    DenseOperands op;  // A = X (K-major for this product), B = R; M = P, N = C, K = observations
    double* Gp;        // [splits][M][ldg]
    long long ldg;
    long long k_per_split;
};

__global__ void __launch_bounds__(DgBwd::THREADS, 1)
dense_bwd_kernel(const __grid_constant__ DenseBwdParams P) {
    using Cfg = DgBwd;
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3, fr = lane >> 2, fk = lane & 3;
    const long long n0 = (long long)blockIdx.x * DG_BN, m0 = (long long)blockIdx.y * DG_BM;
    const long long kb = (long long)blockIdx.z * P.k_per_split;
    const long long ke = kb + P.k_per_split < P.op.K ? kb + P.k_per_split : P.op.K;
    double acc[DG_TM][DG_TN][2];
#pragma unroll
    for (int tm = 0; tm < DG_TM; ++tm)
#pragma unroll
        for (int tn = 0; tn < DG_TN; ++tn) acc[tm][tn][0] = acc[tm][tn][1] = 0.0;
    if (kb < ke) dg_mainloop<false, Cfg>(P.op, smem, m0, n0, kb, ke, acc);
    double* out = P.Gp + (long long)blockIdx.z * P.op.M * P.ldg;
#pragma unroll
    for (int tm = 0; tm < DG_TM; ++tm) {
        const long long p = m0 + wm * 32 + tm * 8 + fr;
        if (p >= P.op.M) continue;
#pragma unroll
        for (int tn = 0; tn < DG_TN; ++tn) {
            const long long c = n0 + wn * 32 + tn * 8 + 2 * fk;
            if (c + 1 < P.op.N) *reinterpret_cast<double2*>(out + p * P.ldg + c) = make_double2(acc[tm][tn][0], acc[tm][tn][1]);
            else if (c < P.op.N) out[p * P.ldg + c] = acc[tm][tn][0];
        }
    }
}

// grad[beta_j][c] += sum over splits (ascending) of Gp[ks][j][c]   (the prior part was written by the plate kernel)
__global__ void __launch_bounds__(BLOCK)
dense_grad_reduce_kernel(const double* __restrict__ Gp, int splits, long long P, long long ldg, double* __restrict__ grad,
                         long long ld, long long theta_base, long long theta_stride, long long C, int accumulate) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    const long long j = blockIdx.y;
    if (c >= C) return;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += Gp[((long long)k * P + j) * ldg + c];
    double* g = grad + (theta_base + theta_stride * j) * ld + c;
    *g = accumulate ? *g + s : s;
}

}  // namespace jbugs
