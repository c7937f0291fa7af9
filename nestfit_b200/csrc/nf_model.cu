// Fused model-synthesis + radiative-transfer + chi-square kernels (sm_100a).
//
// One warp scores one parameter vector against one pixel:
//   lanes <-> hyperfine lines during the FP64 set-up phases,
//   lanes <-> channels (32-channel chunks) during the FP32 main loop.
// A CTA (8 warps) works on a tile of consecutive vectors; the pixel of the
// tile's first vector is staged once in shared memory with a TMA bulk copy
// (cp.async.bulk + mbarrier) and reused by every vector of that pixel.
//
// Reference arithmetic restated here (paths relative to the reference tree):
//   c_amm_predict        nestfit/models/ammonia.pyx:326-361
//   c_partition_level/func                ammonia.pyx:289-315
//   c_hf_predict         nestfit/models/hyperfine.pyx:52-113 (__APPROX window rule 76-96)
//   c_iemtex_interp                       hyperfine.pyx:12-45
//   c_gauss_predict      nestfit/models/gaussian.pyx:17-50
//   Spectrum.c_loglikelihood nestfit/core/core.pyx:522-530
//   FastExp semantics    nestfit/core/fastexp.c:234-283 (exp(-x); 0 for x >= 32;
//                        Taylor-3 below 2^-5) -> MUFU.EX2 with log2(e) folded in.

#include <cmath>
#include <cstdio>
#include <vector>

#include "nf_internal.cuh"
#include "../../include/nf_nh3_tables.h"

#define NF_FULL 0xffffffffu
#define NF_LOG2E 1.4426950408889634
#define NF_HK (NF_H / NF_KB)

// ---- device tables -------------------------------------------------------
__device__ double g_line_freq[NF_NH3_NLINES_TOTAL];  // (1 - voff_i/c) * nu0   hyperfine.pyx:70
__device__ double g_line_wt[NF_NH3_NLINES_TOTAL];    // tau weights            ammonia.pyx:168-228
__device__ double g_iem_y[NF_IEM_SIZE];              // 1/(exp(x_k)-1)         hyperfine.pyx:19
__constant__ double c_iem_xmin, c_iem_xmax, c_iem_step, c_iem_inv_dx;

static const double h_nu[NF_NH3_NTRANS] = NF_NH3_REST_FREQ_INIT;
static const int h_off[NF_NH3_NTRANS + 1] = NF_NH3_LINE_OFFSET_INIT;
static const double h_voff[NF_NH3_NLINES_TOTAL] = NF_NH3_LINE_VOFF_INIT;
static const double h_wt[NF_NH3_NLINES_TOTAL] = NF_NH3_LINE_WEIGHT_INIT;

cudaError_t nf_model_init_device_tables(int device)
{
    static bool done[64] = {false};
    if (device >= 0 && device < 64 && done[device]) return cudaSuccess;
    double freq[NF_NH3_NLINES_TOTAL];
    for (int t = 0; t < NF_NH3_NTRANS; ++t)
        for (int i = h_off[t]; i < h_off[t + 1]; ++i)
            freq[i] = (1.0 - h_voff[i] / NF_CKMS) * h_nu[t];
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(g_line_freq, freq, sizeof(freq)))) return e;
    if ((e = cudaMemcpyToSymbol(g_line_wt, h_wt, sizeof(h_wt)))) return e;
    // hyperfine.pyx:12-20: x = linspace(XMIN, XMAX, 1000), y = 1/(exp(x)-1)
    std::vector<double> y(NF_IEM_SIZE);
    const double lo = NF_H * 23.0e9 / NF_KB, hi = NF_H * 28.0e9 / NF_KB;
    const double xmin = lo / 8.0, xmax = hi / 2.7;
    const double step = (xmax - xmin) / (double)(NF_IEM_SIZE - 1);
    double x1 = xmin + step, inv_dx = 1.0 / (x1 - xmin);
    for (int k = 0; k < NF_IEM_SIZE; ++k) {
        double x = (k == NF_IEM_SIZE - 1) ? xmax : xmin + (double)k * step;
        y[k] = 1.0 / (std::exp(x) - 1.0);
    }
    if ((e = cudaMemcpyToSymbol(g_iem_y, y.data(), sizeof(double) * NF_IEM_SIZE))) return e;
    if ((e = cudaMemcpyToSymbol(c_iem_xmin, &xmin, sizeof(double)))) return e;
    if ((e = cudaMemcpyToSymbol(c_iem_xmax, &xmax, sizeof(double)))) return e;
    if ((e = cudaMemcpyToSymbol(c_iem_step, &step, sizeof(double)))) return e;
    if ((e = cudaMemcpyToSymbol(c_iem_inv_dx, &inv_dx, sizeof(double)))) return e;
    if (device >= 0 && device < 64) done[device] = true;
    return cudaSuccess;
}

// ---- small device helpers -------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NF_FULL, v, o);
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// 1/(exp(x)-1) exactly as the reference evaluates it (table lerp inside the
// table domain, expm1 outside), hyperfine.pyx:23-45.  FP64.
__device__ double iemtex_ref(double x)
{
    if (c_iem_xmin < x && x < c_iem_xmax) {
        int k = (int)((x - c_iem_xmin) * c_iem_inv_dx);
        k = min(k, NF_IEM_SIZE - 2);
        double xk = (k == NF_IEM_SIZE - 1) ? c_iem_xmax : c_iem_xmin + (double)k * c_iem_step;
        double yk = g_iem_y[k], yk1 = g_iem_y[k + 1];
        return (yk1 - yk) * c_iem_inv_dx * (x - xk) + yk;
    }
    return 1.0 / expm1(x);
}

// exp(-x) with FastExp's argument handling (float-rounded argument, zero from 32 up)
__device__ __forceinline__ double fastexp_f64(double x)
{
    float xf = (float)x;
    return (xf < 32.0f) ? exp(-(double)xf) : 0.0;
}

template <typename T>
__device__ __forceinline__ double ld_param(const void *base, int64_t idx)
{
    return (double)__ldg(reinterpret_cast<const T *>(base) + idx);
}

// Per-warp scratch in shared memory.
template <int NC>
struct __align__(16) WarpScratch {
    float4 lineA[NC][NF_MAX_LINES];        // {R, -k2, 2*k2*phi, -k2*phi^2}
    float4 lineB[NC][NF_MAX_LINES];        // {tau weight, lo - R, hi - R, 0}
    float4 amp[NC][NF_MAX_SPEC];           // {aL, bL, aR, bR} of T_B amplitude lines
    double tauT[NC][NF_MAX_SPEC];          // main-line optical depth
    double soc[NC], voc[NC];               // sigma / c_kms, voff / c_kms
};

// ---- the fused kernel -----------------------------------------------------
template <int NC, bool IS_NH3, bool WRITE_PRED, typename PT>
__global__ void __launch_bounds__(NF_THREADS)
nf_like_kernel(const __grid_constant__ NfLikeArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float *sdata = reinterpret_cast<float *>(smem_raw + 128);
    const int data_floats = a.n_spec * a.n_pad;
    WarpScratch<NC> *scr_all =
        reinterpret_cast<WarpScratch<NC> *>(smem_raw + 128 + (((size_t)data_floats * 4 + 127) / 128) * 128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpScratch<NC> &sc = scr_all[warp];

    const int64_t b0 = (int64_t)blockIdx.x * NF_TILE_VECS;
    const bool have_data = a.data != nullptr;
    int64_t pix0 = 0;
    if (have_data) {
        pix0 = a.pix_of_vec ? (int64_t)__ldg(a.pix_of_vec + b0) : b0 / a.vecs_per_pix;
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0)
            tma_load_1d(sdata, a.data + pix0 * a.pix_stride, (uint32_t)(data_floats * 4), bar);
    }
    bool data_ready = !have_data;

    const int ncomp = IS_NH3 ? NC : a.ncomp;   // NH3: template; Gaussian: lines of one group
    const int ndim = IS_NH3 ? 6 * NC : 3 * a.ncomp;
    const int nchunks = (a.n_chan + 31) >> 5;
    const float xl = (float)lane;

    for (int64_t b = b0 + warp; b < b0 + NF_TILE_VECS && b < a.B; b += NF_WARPS_PER_CTA) {
        const int64_t pbase = b * ndim;
        int64_t pix = 0;
        if (have_data) pix = a.pix_of_vec ? (int64_t)__ldg(a.pix_of_vec + b) : b / a.vecs_per_pix;

        if (IS_NH3) {
            // ---- P1: partition function, lanes <-> J (ammonia.pyx:289-315) ----
            double zlev = 0.0, qtot = 1.0, trot_mine = 1.0;
            const int my_c = lane / a.n_spec, my_s = lane - my_c * a.n_spec;
            const bool pair_lane = lane < NC * a.n_spec;
            const int my_J = pair_lane ? a.spec[my_s].J : 0;
            const int my_para = pair_lane ? a.spec[my_s].para : 1;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                double trot = ld_param<PT>(a.params, pbase + 1 * NC + c);
                if (a.cold)  // swift_convert, ammonia.pyx:280-286
                    trot = trot / (1.0 + (trot / 41.18) * log(1.0 + 0.6 * exp(-15.7 / trot)));
                const double J = (double)lane;
                double aJ = NF_HK * (NF_BROT * J * (J + 1.0) + (NF_CROT - NF_BROT) * J * J);
                double lev = (2.0 * J + 1.0) * fastexp_f64(aJ / trot);
                double qp = (lane % 3 != 0) ? lev : 0.0;
                double qo = (lane % 3 == 0) ? 2.0 * lev : 0.0;
                if (!(trot < 299.0)) {  // levels J >= 32 underflow FastExp's range below ~301 K
                    const double J2 = (double)(lane + 32);
                    if (lane + 32 <= 50) {
                        double a2 = NF_HK * (NF_BROT * J2 * (J2 + 1.0) + (NF_CROT - NF_BROT) * J2 * J2);
                        double l2 = (2.0 * J2 + 1.0) * fastexp_f64(a2 / trot);
                        if ((lane + 32) % 3 != 0) qp += l2; else qo += 2.0 * l2;
                    }
                }
                double zl = __shfl_sync(NF_FULL, lev, my_J);
                if (a.need_para) qp = warp_sum(qp);
                if (a.need_ortho) qo = warp_sum(qo);
                if (my_c == c) { zlev = zl; qtot = my_para ? qp : qo; trot_mine = trot; }
            }
            // ---- P2: per (component, spectrum) scalars, lanes <-> pairs ----
            if (pair_lane) {
                const NfSpecMeta &sm = a.spec[my_s];
                const double voff = ld_param<PT>(a.params, pbase + 0 * NC + my_c);
                double tex = ld_param<PT>(a.params, pbase + 2 * NC + my_c);
                const double ntot = ld_param<PT>(a.params, pbase + 3 * NC + my_c);
                const double sigm = ld_param<PT>(a.params, pbase + 4 * NC + my_c);
                const double orth = ld_param<PT>(a.params, pbase + 5 * NC + my_c);
                if (a.lte) tex = trot_mine;
                const double frac = my_para ? 1.0 - orth : orth;
                const double pop = exp10(ntot) * frac * zlev / qtot;          // ammonia.pyx:353
                const double e = exp(-sm.hnu_k / tex);                        // ammonia.pyx:354-357
                const double tau_main = pop * sm.fracterm * ((1.0 - e) / (1.0 + e)) * (sm.width_c / sigm);
                sc.tauT[my_c][my_s] = tau_main;
                if (my_s == 0) { sc.soc[my_c] = sigm / NF_CKMS; sc.voc[my_c] = voff / NF_CKMS; }
                // T_B amplitude  T0_j * (G(T0_j/tex) - tbg_j), hyperfine.pyx:106-113, as the max
                // of two lines in j (the reference's G is a convex piecewise-linear table).
                const double nm1 = (double)(a.n_chan - 1);
                const double xL = sm.T0_first / tex, xR = sm.T0_last / tex;
                const double dxdj = (xR - xL) / nm1;
                double aL, bL, aR, bR;
                const bool inL = c_iem_xmin < xL && xL < c_iem_xmax;
                const bool inR = c_iem_xmin < xR && xR < c_iem_xmax;
                if (inL && inR) {
                    int kL = min((int)((xL - c_iem_xmin) * c_iem_inv_dx), NF_IEM_SIZE - 2);
                    int kR = min((int)((xR - c_iem_xmin) * c_iem_inv_dx), NF_IEM_SIZE - 2);
                    double xk = c_iem_xmin + (double)kL * c_iem_step;
                    double sl = (g_iem_y[kL + 1] - g_iem_y[kL]) * c_iem_inv_dx;
                    aL = g_iem_y[kL] + sl * (xL - xk);
                    bL = sl * dxdj;
                    xk = c_iem_xmin + (double)kR * c_iem_step;
                    sl = (g_iem_y[kR + 1] - g_iem_y[kR]) * c_iem_inv_dx;
                    aR = g_iem_y[kR] + sl * (xL - xk);
                    bR = sl * dxdj;
                } else {
                    const double gL = iemtex_ref(xL), gR = iemtex_ref(xR);
                    aL = aR = gL;
                    bL = bR = (gR - gL) / nm1;
                }
                sc.amp[my_c][my_s] = make_float4((float)(aL - sm.tbg0), (float)(bL - sm.tbg1),
                                                 (float)(aR - sm.tbg0), (float)(bR - sm.tbg1));
            }
            __syncwarp();
        }

        double lnl = 0.0;
        for (int s = 0; s < a.n_spec; ++s) {
            const NfSpecMeta &sm = a.spec[s];
            // ---- P3: per-line window + Gaussian coefficients, lanes <-> lines ----
            uint32_t lohi[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int nl = IS_NH3 ? sm.nlines : ncomp;
                const bool act = lane < nl;
                double f, soc, voc, wT, nucen, w;
                if (IS_NH3) {
                    f = act ? g_line_freq[sm.line_off + lane] : sm.nu0;
                    soc = sc.soc[c]; voc = sc.voc[c];
                    wT = act ? sc.tauT[c][s] * g_line_wt[sm.line_off + lane] : 0.0;
                    w = soc * f;                  // hyperfine.pyx:71
                    nucen = f - voc * f;          // hyperfine.pyx:72-73
                } else {
                    f = sm.nu0;                   // gaussian.pyx:28-33
                    const int cl = act ? lane : 0;
                    const double voff = ld_param<PT>(a.params, pbase + cl);
                    const double sigm = ld_param<PT>(a.params, pbase + ncomp + cl);
                    wT = act ? ld_param<PT>(a.params, pbase + 2 * ncomp + cl) : 0.0;
                    w = sigm / NF_CKMS * f;
                    nucen = f * (1.0 - voff / NF_CKMS);
                }
                const double cut = 5.0 * fabs(w);          // sqrt(12.5 / (0.5 / w^2)), hyperfine.pyx:82
                const double rel = nucen - sm.nu_min;
                const double nmax = (double)a.n_chan;
                // floor((nu_cen - nu_min -/+ cut) / nu_chan), hyperfine.pyx:83-87
                double flo = floor((rel - cut) / sm.nu_chan), fhi = floor((rel + cut) / sm.nu_chan);
                flo = fmax(fmin(flo, nmax), -1.0);          // NaN -> out of range -> skipped
                fhi = fmax(fmin(fhi, nmax), -1.0);
                if (!(flo == flo) || !(fhi == fhi)) { flo = nmax; fhi = -1.0; }
                int lo = (int)flo, hi = (int)fhi;
                bool on = act && !(hi < 0 || lo > a.n_chan - 1);   // hyperfine.pyx:88
                lo = max(lo, 0);
                hi = min(hi, a.n_chan - 1);
                on = on && hi > lo;                                // loop j in [lo, hi)
                const double jc = rel * sm.inv_chan;
                double R = rint(jc);
                R = fmax(fmin(R, 1.0e7), -1.0e7);
                const double phi = jc - R;
                const double sch = w * sm.inv_chan;
                const double k2 = 0.5 / (sch * sch) * NF_LOG2E;
                float4 A, Bv;
                A.x = (float)R; A.y = (float)(-k2); A.z = (float)(2.0 * k2 * phi); A.w = (float)(-k2 * phi * phi);
                Bv.x = on ? (float)wT : 0.0f;
                Bv.y = on ? (float)((double)lo - R) : 0.0f;
                Bv.z = on ? (float)((double)hi - R) : 0.0f;
                Bv.w = 0.0f;
                if (!on) { A.x = 0.f; A.y = 0.f; A.z = 0.f; A.w = 0.f; }
                sc.lineA[c][lane] = A;
                sc.lineB[c][lane] = Bv;
                lohi[c] = on ? ((uint32_t)lo | ((uint32_t)hi << 16)) : 0u;
            }
            __syncwarp();
            if (!data_ready) { mbar_wait(bar, 0); data_ready = true; }

            const bool staged = have_data && pix == pix0;
            const float *grow = have_data ? a.data + pix * a.pix_stride + (int64_t)s * a.n_pad : nullptr;
            const float *srow = sdata + s * a.n_pad;
            float acc = 0.0f;
            for (int sb = 0; sb < nchunks; sb += 32) {
                uint32_t cm[NC], un[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    // chunks [sb, sb+32) touched by this lane's line
                    int lo = (int)(lohi[c] & 0xffffu), hi = (int)(lohi[c] >> 16);
                    int c_lo = (lo >> 5) - sb, c_hi = ((hi - 1) >> 5) - sb;
                    uint32_t m = 0u;
                    if (hi > lo && c_hi >= 0 && c_lo < 32) {
                        c_lo = max(c_lo, 0); c_hi = min(c_hi, 31);
                        m = (0xffffffffu >> (31 - c_hi)) & (0xffffffffu << c_lo);
                    }
                    cm[c] = m;
                    un[c] = __reduce_or_sync(NF_FULL, m);
                }
                const int cend = min(32, nchunks - sb);
                for (int cc = 0; cc < cend; ++cc) {
                    const int j = ((sb + cc) << 5) + lane;
                    const float xj = (float)((sb + cc) << 5) + xl;
                    float d = 0.0f;
                    if (have_data) d = staged ? srow[j] : __ldg(grow + j);
                    float m = 0.0f;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        if (!((un[c] >> cc) & 1u)) continue;
                        uint32_t lm = __ballot_sync(NF_FULL, (cm[c] >> cc) & 1u);
                        float tau = 0.0f;
                        while (lm) {
                            const int i = __ffs(lm) - 1;
                            lm &= lm - 1;
                            const float4 A = sc.lineA[c][i];
                            const float4 Bv = sc.lineB[c][i];
                            const float d0 = xj - A.x;
                            const float t = fmaf(A.y, d0, A.z);
                            const float e = ex2_approx(fmaf(t, d0, A.w));
                            if (d0 >= Bv.y && d0 < Bv.z) tau = fmaf(Bv.x, e, tau);
                        }
                        if (IS_NH3) {
                            const float4 am = sc.amp[c][s];
                            const float T0 = fmaf(sm.t0b, xj, sm.t0a);
                            const float D = fmaxf(fmaf(am.y, xj, am.x), fmaf(am.w, xj, am.z));
                            float e1;
                            if (fabsf(tau) < 0.03125f)   // FastExp Taylor branch, fastexp.c:265-270
                                e1 = tau * (1.0f - 0.5f * tau * (1.0f - tau * (1.0f / 3.0f)));
                            else
                                e1 = 1.0f - ex2_approx(-(float)NF_LOG2E * tau);
                            m = fmaf(T0 * D, e1, m);
                        } else {
                            m += tau;
                        }
                    }
                    if (WRITE_PRED) {
                        if (j < a.n_chan) a.pred[(b * a.n_spec + s) * (int64_t)a.n_chan + j] = m;
                    }
                    const float r = d - m;
                    acc = fmaf(r, r, acc);
                }
            }
            if (have_data) {
                const double tot = warp_sum((double)acc);
                lnl -= tot * __ldg(a.inv2s2 + pix * a.n_spec + s);
            }
            __syncwarp();
        }
        if (a.lnL && lane == 0) a.lnL[b] = lnl;
    }
    // a CTA whose warps all ran out of vectors must still drain the bulk copy
    if (!data_ready) mbar_wait(bar, 0);
}

template <int NC>
static size_t like_smem_bytes(const NfLikeArgs &a)
{
    size_t data = (((size_t)a.n_spec * a.n_pad * 4 + 127) / 128) * 128;
    return 128 + data + sizeof(WarpScratch<NC>) * NF_WARPS_PER_CTA;
}

template <int NC, bool IS_NH3, bool WP, typename PT>
static cudaError_t launch_one(const NfLikeArgs &a, cudaStream_t st)
{
    auto kern = nf_like_kernel<NC, IS_NH3, WP, PT>;
    size_t smem = like_smem_bytes<NC>(a);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    int64_t grid = (a.B + NF_TILE_VECS - 1) / NF_TILE_VECS;
    if (grid <= 0) return cudaSuccess;
    kern<<<(unsigned)grid, NF_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int NC, bool IS_NH3>
static cudaError_t launch_nc(const NfLikeArgs &a, cudaStream_t st)
{
    const bool wp = a.pred != nullptr;
    if (a.param_f64)
        return wp ? launch_one<NC, IS_NH3, true, double>(a, st) : launch_one<NC, IS_NH3, false, double>(a, st);
    return wp ? launch_one<NC, IS_NH3, true, float>(a, st) : launch_one<NC, IS_NH3, false, float>(a, st);
}

cudaError_t nf_launch_nh3(const NfLikeArgs &a, cudaStream_t st)
{
    switch (a.ncomp) {
    case 1: return launch_nc<1, true>(a, st);
    case 2: return launch_nc<2, true>(a, st);
    case 3: return launch_nc<3, true>(a, st);
    case 4: return launch_nc<4, true>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t nf_launch_gauss(const NfLikeArgs &a, cudaStream_t st)
{
    return launch_nc<1, false>(a, st);
}

// ---- pixel-block helpers ----------------------------------------------------
// null_lnZ = -sum_s sum_j d^2 / (2 sigma_s^2)   (core.pyx:518-520, ammonia.pyx:411-415)
__global__ void nf_null_lnz_kernel(const float *data, const double *inv2s2, double *out,
                                   int64_t n_pix, int n_spec, int n_chan, int n_pad)
{
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_pix) return;
    double tot = 0.0;
    for (int s = 0; s < n_spec; ++s) {
        const float *row = data + (p * n_spec + s) * (int64_t)n_pad;
        double acc = 0.0;
        for (int j = lane; j < n_chan; j += 32) { double d = (double)row[j]; acc += d * d; }
        acc = warp_sum(acc);
        tot -= acc * inv2s2[p * n_spec + s];
    }
    if (lane == 0) out[p] = tot;
}

cudaError_t nf_launch_null_lnz(const float *data, const double *inv2s2, double *out, int64_t n_pix,
                               int n_spec, int n_chan, int n_pad, cudaStream_t st)
{
    const int wpb = 8;
    int64_t grid = (n_pix + wpb - 1) / wpb;
    if (grid <= 0) return cudaSuccess;
    nf_null_lnz_kernel<<<(unsigned)grid, wpb * 32, 0, st>>>(data, inv2s2, out, n_pix, n_spec, n_chan, n_pad);
    return cudaGetLastError();
}

// rows of n_chan (f32 or f64) -> zero-padded FP32 rows of n_pad
template <typename T>
__global__ void nf_pack_rows_kernel(const T *src, float *dst, int64_t rows, int n_chan, int n_pad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = rows * n_pad;
    for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / n_pad;
        int j = (int)(i - r * n_pad);
        dst[i] = j < n_chan ? (float)src[r * n_chan + j] : 0.0f;
    }
}

cudaError_t nf_launch_pack_rows(const void *src, int src_f64, float *dst, int64_t rows, int n_chan,
                                int n_pad, cudaStream_t st)
{
    int64_t total = rows * n_pad;
    if (total <= 0) return cudaSuccess;
    int64_t grid = (total + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    if (src_f64)
        nf_pack_rows_kernel<double><<<(unsigned)grid, 256, 0, st>>>((const double *)src, dst, rows, n_chan, n_pad);
    else
        nf_pack_rows_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float *)src, dst, rows, n_chan, n_pad);
    return cudaGetLastError();
}
