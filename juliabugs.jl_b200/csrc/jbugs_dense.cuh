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
// tcgen05 has no FP64 kind).  CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 MMA tiles, three
// cp.async stages.  Shared-memory rows are padded to a stride == 4 (mod 16) doubles so that the 64-bit fragment
// loads of a half-warp hit 16 distinct bank pairs.
#pragma once

#include "jbugs_kernels.cuh"

namespace jbugs {

constexpr int DG_BM = 128, DG_BN = 128, DG_BK = 16, DG_STAGES = 3, DG_THREADS = 256;
constexpr int DG_LDB = DG_BN + 4;   // Bs[k][n]
constexpr int DG_LDA_M = DG_BM + 4; // As[k][m]   (A is M-major in memory: forward GEMM)
constexpr int DG_LDA_K = DG_BK + 4; // As[m][k]   (A is K-major in memory: reverse GEMM)
constexpr int DG_A_DOUBLES = (DG_BK * DG_LDA_M > DG_BM * DG_LDA_K) ? DG_BK * DG_LDA_M : DG_BM * DG_LDA_K;
constexpr int DG_STAGE_DOUBLES = DG_A_DOUBLES + DG_BK * DG_LDB;
constexpr size_t DG_SMEM_BYTES = (size_t)DG_STAGES * DG_STAGE_DOUBLES * sizeof(double);

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

// 8-byte cp.async with zero fill when `ok` is false
__device__ __forceinline__ void cp_async8_zfill(double* smem_dst, const double* gmem_src, bool ok) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

struct DenseOperands {
    const double* A;  // X
    long long lda;    // number of observations (leading dimension of the column-major X)
    const double* B;  // forward: theta + beta_base*ld (row j at j*ldb_rows); reverse: R
    long long ldb;    // elements between consecutive k rows of B
    long long M, N, K;
};

// load one BK-slab of A and B into a stage.  A_MMAJOR: A(m,k) = A[m + lda*k]; else A(m,k) = A[k + lda*m]
template <bool A_MMAJOR>
__device__ __forceinline__ void dg_load_stage(const DenseOperands& op, double* stage, long long m0, long long n0,
                                              long long k0, int tid) {
    double* As = stage;
    double* Bs = stage + DG_A_DOUBLES;
    if (A_MMAJOR) {
        // As[kk][m]: consecutive threads walk m (contiguous in memory)
        for (int q = tid; q < DG_BK * DG_BM; q += DG_THREADS) {
            const int kk = q / DG_BM, m = q % DG_BM;
            const bool ok = (m0 + m < op.M) && (k0 + kk < op.K);
            cp_async8_zfill(As + kk * DG_LDA_M + m, op.A + (ok ? (m0 + m) + op.lda * (k0 + kk) : 0), ok);
        }
    } else {
        // As[m][kk]: consecutive threads walk kk (contiguous in memory)
        for (int q = tid; q < DG_BK * DG_BM; q += DG_THREADS) {
            const int m = q / DG_BK, kk = q % DG_BK;
            const bool ok = (m0 + m < op.M) && (k0 + kk < op.K);
            cp_async8_zfill(As + m * DG_LDA_K + kk, op.A + (ok ? (k0 + kk) + op.lda * (m0 + m) : 0), ok);
        }
    }
    for (int q = tid; q < DG_BK * DG_BN; q += DG_THREADS) {
        const int kk = q / DG_BN, n = q % DG_BN;
        const bool ok = (n0 + n < op.N) && (k0 + kk < op.K);
        cp_async8_zfill(Bs + kk * DG_LDB + n, op.B + (ok ? (k0 + kk) * op.ldb + (n0 + n) : 0), ok);
    }
}

// acc[tm][tn][2] += A_tile * B_tile over k in [k_begin, k_end)
template <bool A_MMAJOR>
__device__ __forceinline__ void dg_mainloop(const DenseOperands& op, double* smem, long long m0, long long n0,
                                            long long k_begin, long long k_end, double (&acc)[8][4][2]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index
    const long long nk = (k_end - k_begin + DG_BK - 1) / DG_BK;
    // prologue: stages 0 .. STAGES-2
    for (int s = 0; s < DG_STAGES - 1; ++s) {
        if (s < nk) dg_load_stage<A_MMAJOR>(op, smem + s * DG_STAGE_DOUBLES, m0, n0, k_begin + (long long)s * DG_BK, tid);
        cp_async_commit();
    }
    for (long long kt = 0; kt < nk; ++kt) {
        cp_async_wait_1();
        __syncthreads();
        // prefetch slab kt + STAGES - 1 into the stage consumed at iteration kt - 1
        const long long pf = kt + DG_STAGES - 1;
        if (pf < nk) dg_load_stage<A_MMAJOR>(op, smem + (pf % DG_STAGES) * DG_STAGE_DOUBLES, m0, n0, k_begin + pf * DG_BK, tid);
        cp_async_commit();
        const double* As = smem + (kt % DG_STAGES) * DG_STAGE_DOUBLES;
        const double* Bs = As + DG_A_DOUBLES;
#pragma unroll
        for (int ks = 0; ks < DG_BK / 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int tm = 0; tm < 8; ++tm) {
                const int m = wm * 64 + tm * 8 + fr, k = ks * 4 + fk;
                a[tm] = A_MMAJOR ? As[k * DG_LDA_M + m] : As[m * DG_LDA_K + k];
            }
#pragma unroll
            for (int tn = 0; tn < 4; ++tn) b[tn] = Bs[(ks * 4 + fk) * DG_LDB + wn * 32 + tn * 8 + fr];
#pragma unroll
            for (int tm = 0; tm < 8; ++tm)
#pragma unroll
                for (int tn = 0; tn < 4; ++tn) dmma(acc[tm][tn], a[tm], b[tn]);
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

__global__ void __launch_bounds__(DG_THREADS, 1)
dense_fwd_kernel(const __grid_constant__ DenseFwdParams P) {
    extern __shared__ double smem[];
    __shared__ double red[2][DG_BN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3, fr = lane >> 2, fk = lane & 3;
    const long long n0 = (long long)blockIdx.x * DG_BN;
    const long long m_tiles = (P.op.M + DG_BM - 1) / DG_BM;
    double lp[4][2];  // per-thread log-density sums for its 4 x 2 chain columns
#pragma unroll
    for (int tn = 0; tn < 4; ++tn) lp[tn][0] = lp[tn][1] = 0.0;
    for (long long mt = blockIdx.y; mt < m_tiles; mt += gridDim.y) {
        const long long m0 = mt * DG_BM;
        double acc[8][4][2];
#pragma unroll
        for (int tm = 0; tm < 8; ++tm)
#pragma unroll
            for (int tn = 0; tn < 4; ++tn) acc[tm][tn][0] = acc[tm][tn][1] = 0.0;
        dg_mainloop<true>(P.op, smem, m0, n0, 0, P.op.K, acc);
        // epilogue: C fragment element (row = fr, col = 2*fk + {0,1}) of MMA tile (tm, tn)
#pragma unroll
        for (int tm = 0; tm < 8; ++tm) {
            const long long i = m0 + wm * 64 + tm * 8 + fr;
            const bool row_ok = i < P.op.M;
            const double yv = row_ok ? __ldg(P.y + i) : 0.0;
            const double nv = row_ok ? (P.n ? __ldg(P.n + i) : 1.0) : 0.0;
            const double mv = nv - yv;
#pragma unroll
            for (int tn = 0; tn < 4; ++tn) {
                double r2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double eta = acc[tm][tn][e];
                    const double ax = fabs(eta);
                    const double ex = exp(-ax);
                    const double l1p = log1p(ex);
                    const double inv = 1.0 / (1.0 + ex);
                    const double p = eta >= 0.0 ? inv : ex * inv;
                    const double lsp = -((eta >= 0.0 ? 0.0 : ax) + l1p);
                    const double lsq = -((eta >= 0.0 ? ax : 0.0) + l1p);
                    double term = (yv == 0.0 ? 0.0 : yv * lsp) + (mv == 0.0 ? 0.0 : mv * lsq);
                    if (eta > kLogisticUpper && mv != 0.0) term = -dev_inf();
                    if (eta < kLogisticLower && yv != 0.0) term = -dev_inf();
                    lp[tn][e] += row_ok ? term : 0.0;
                    r2[e] = yv - nv * p;
                }
                const long long c = n0 + wn * 32 + tn * 8 + 2 * fk;
                if (row_ok && c < P.op.N) {
                    if (c + 1 < P.op.N) *reinterpret_cast<double2*>(P.R + i * P.ldr + c) = make_double2(r2[0], r2[1]);
                    else P.R[i * P.ldr + c] = r2[0];
                }
            }
        }
    }
    // fixed-order reduction of lp over the 8 fragment rows (shuffles), then over the two M-warps (shared memory)
#pragma unroll
    for (int tn = 0; tn < 4; ++tn)
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
        if (c < P.op.N) P.partial[(long long)blockIdx.y * P.ldg + c] = red[0][tid] + red[1][tid];
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

__global__ void __launch_bounds__(DG_THREADS, 1)
dense_bwd_kernel(const __grid_constant__ DenseBwdParams P) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3, fr = lane >> 2, fk = lane & 3;
    const long long n0 = (long long)blockIdx.x * DG_BN, m0 = (long long)blockIdx.y * DG_BM;
    const long long kb = (long long)blockIdx.z * P.k_per_split;
    const long long ke = kb + P.k_per_split < P.op.K ? kb + P.k_per_split : P.op.K;
    double acc[8][4][2];
#pragma unroll
    for (int tm = 0; tm < 8; ++tm)
#pragma unroll
        for (int tn = 0; tn < 4; ++tn) acc[tm][tn][0] = acc[tm][tn][1] = 0.0;
    if (kb < ke) dg_mainloop<false>(P.op, smem, m0, n0, kb, ke, acc);
    double* out = P.Gp + (long long)blockIdx.z * P.op.M * P.ldg;
#pragma unroll
    for (int tm = 0; tm < 8; ++tm) {
        const long long p = m0 + wm * 64 + tm * 8 + fr;
        if (p >= P.op.M) continue;
#pragma unroll
        for (int tn = 0; tn < 4; ++tn) {
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
