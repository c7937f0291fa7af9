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
#include "nf_device.cuh"
#include "../../include/nf_nh3_tables.h"

#ifndef NF_MIN_CTAS
#define NF_MIN_CTAS 3
#endif

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


// One hyperfine line of one (component, spectrum): 32 bytes, read with one LDS.128 + one LDS.64.
struct __align__(32) LineRec {
    float4 a;   // {R, -k2, 2*k2*phi, weight * 2^(-k2*phi^2)}
    float2 w;   // {lo - R, hi - R}: window [lo, hi) in the line's own integer frame
    float2 pad;
};

// Two adjacent hyperfine lines, element-interleaved so that the packed FP32x2 pipeline
// (FADD2 / FFMA2 / FMUL2) works on both at once: 48 bytes = three LDS.128.
struct __align__(16) PairRec {
    float4 a;   // {R0, R1, -k2_0, -k2_1}
    float4 b;   // {2 k2 phi (0), (1), weight*2^(-k2 phi^2) (0), (1)}
    float4 w;   // {lo-R (0), (1), hi-R (0), (1)}
};
#define NF_MAX_PAIRS 17     // pairs (2q+p, 2q+p+1) of up to 33 records (32 lines + a null line)

// Per-warp scratch in shared memory.
template <int NC, bool NH3>
struct __align__(32) WarpScratch;

template <int NC>
struct __align__(32) WarpScratch<NC, true> {
    // pair[c][p][q] holds lines (2q+p, 2q+p+1): a run of lines may start at either parity
    PairRec pair[NC][2][NF_MAX_PAIRS];
    uint4 tab[32];                         // per chunk: (first pair offset | pair count << 16) per component
    float4 amp[NC][NF_MAX_SPEC];           // {aL, bL, aR, bR} of T_B amplitude lines
    double tauT[NC][NF_MAX_SPEC];          // main-line optical depth
    double soc[NC], voc[NC];               // sigma / c_kms, voff / c_kms
};

template <int NC>
struct __align__(32) WarpScratch<NC, false> {
    LineRec line[NC][NF_MAX_LINES];
    float4 amp[NC][NF_MAX_SPEC];
    double tauT[NC][NF_MAX_SPEC];
    double soc[NC], voc[NC];
};


__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// tau += w * e  for lanes whose channel lies inside the line's window [dlo, dhi)
__device__ __forceinline__ void masked_fma(float &tau, float w, float e, float d0, float dlo, float dhi)
{
    asm("{\n"
        ".reg .pred p;\n"
        "setp.ge.f32 p, %1, %2;\n"
        "setp.lt.and.f32 p, %1, %3, p;\n"
        "@p fma.rn.f32 %0, %4, %5, %0;\n"
        "}\n"
        : "+f"(tau)
        : "f"(d0), "f"(dlo), "f"(dhi), "f"(w), "f"(e));
}


// Two windowed Gaussian terms (lines 2q+p and 2q+p+1) at channel coordinate xj (packed twice).
// The record stores -R so that d0 = xj + (-R) is a single FADD2.
__device__ __forceinline__ void pair_term(float &tau0, float &tau1, const PairRec *rec, uint64_t xj2)
{
    const float4 A = rec->a, B = rec->b, W = rec->w;
    const uint64_t d2 = add2(xj2, pack2(A.x, A.y));                       // exact: integer-valued floats
    const uint64_t t2 = fma2(pack2(A.z, A.w), d2, pack2(B.x, B.y));
    const uint64_t a2 = mul2(t2, d2);
    float d0, d1, a0, a1;
    unpack2(d2, d0, d1);
    unpack2(a2, a0, a1);
    const float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
    masked_fma(tau0, B.z, e0, d0, W.x, W.z);
    masked_fma(tau1, B.w, e1, d1, W.y, W.w);
}

// One windowed Gaussian term of line record (A, Bw) at channel coordinate xj.
__device__ __forceinline__ void line_term(float &tau, const LineRec *rec, float xj)
{
    const float4 A = rec->a;
    const float2 Bw = rec->w;
    const float d0 = xj - A.x;                 // exact: integer-valued floats
    const float t = fmaf(A.y, d0, A.z);
    const float e = ex2_approx(t * d0);        // 2^(-k2 (d0^2 - 2 phi d0)); 2^(-k2 phi^2) is in A.w
    masked_fma(tau, A.w, e, d0, Bw.x, Bw.y);
}

// Number of lanes whose (lane-sorted, ascending) key is <= g; lane 31 must hold a sentinel.
__device__ __forceinline__ int count_le_sorted(int key, int g)
{
    int pos = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(NF_FULL, key, pos + step - 1);
        if (v <= g) pos += step;
    }
    return pos;
}

// (2J+1) * h (B J(J+1) + (C-B) J^2) / k_B in kelvin, FP32 (levels J >= 3 and J = 0)
#define NF_BK_F ((float)(NF_HK * NF_BROT))
#define NF_CK_F ((float)(NF_HK * (NF_CROT - NF_BROT)))

__device__ __forceinline__ float level_f32(int J, float inv_trot)
{
    const float Jf = (float)J;
    const float x = (NF_BK_F * Jf * (Jf + 1.0f) + NF_CK_F * Jf * Jf) * inv_trot;
    return x < 32.0f ? (2.0f * Jf + 1.0f) * ex2_approx(-(float)NF_LOG2E * x) : 0.0f;
}

// ---- the fused kernel -----------------------------------------------------
// WRITE_PRED = false: log-likelihood against the pixel's data (a.data, a.lnL);
// WRITE_PRED = true : model spectra only (a.pred), no data are read.
template <int NC, bool IS_NH3, bool WRITE_PRED, typename PT>
__global__ void __launch_bounds__(NF_THREADS, NF_MIN_CTAS)
nf_like_kernel(const __grid_constant__ NfLikeArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float *sdata = reinterpret_cast<float *>(smem_raw + 128);
    const int data_floats = a.n_spec * a.n_pad;
    typedef WarpScratch<NC, IS_NH3> Scratch;
    Scratch *scr_all = reinterpret_cast<Scratch *>(smem_raw + 128 + (((size_t)data_floats * 4 + 127) / 128) * 128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Scratch &sc = scr_all[warp];

    const int tile = a.tile_vecs > 0 ? a.tile_vecs : NF_TILE_VECS;
    const int64_t b0 = (int64_t)blockIdx.x * tile;
    // the vector count may live on the device (lock-step sampler): a.B then only sized the grid
    const int64_t B = a.B_dev ? min(a.B, (int64_t)__ldg(a.B_dev)) : a.B;
    if (b0 >= B) return;
    constexpr bool have_data = !WRITE_PRED;
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
    const uint32_t sdata_addr = smem_u32(sdata) + (uint32_t)lane * 4u;

    for (int64_t b = b0 + warp; b < b0 + tile && b < B; b += NF_WARPS_PER_CTA) {
        const int64_t pbase = b * ndim;
        int64_t pix = 0;
        if (have_data) pix = a.pix_of_vec ? (int64_t)__ldg(a.pix_of_vec + b) : b / a.vecs_per_pix;

        if (IS_NH3) {
            // ---- P1: partition sums over J, lanes <-> J (ammonia.pyx:289-315).  Levels
            // J = 1, 2 (which carry all but ~1e-3 of Q_para) are added in FP64 in P2; the
            // rest go through MUFU.EX2 in FP32. ----
            float q32 = 0.0f;
            double trot_mine = 1.0;
            const int my_c = lane / a.n_spec, my_s = lane - my_c * a.n_spec;
            const bool pair_lane = lane < NC * a.n_spec;
            const int my_J = pair_lane ? a.spec[my_s].J : 1;
            const int my_para = pair_lane ? a.spec[my_s].para : 1;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                double trot = ld_param<PT>(a.params, pbase + 1 * NC + c);
                if (a.cold)  // swift_convert, ammonia.pyx:280-286
                    trot = trot / (1.0 + (trot / 41.18) * log(1.0 + 0.6 * exp(-15.7 / trot)));
                const float itr = 1.0f / (float)trot;
                float lev = (lane == 1 || lane == 2) ? 0.0f : level_f32(lane, itr);
                float qp = (lane % 3 != 0) ? lev : 0.0f;
                float qo = (lane % 3 == 0) ? 2.0f * lev : 0.0f;
                if (!(trot < 299.0)) {  // levels J >= 32 are zero in FastExp's range below ~301 K
                    if (lane + 32 <= 50) {
                        const float l2 = level_f32(lane + 32, itr);
                        if ((lane + 32) % 3 != 0) qp += l2; else qo += 2.0f * l2;
                    }
                }
                if (a.need_para) qp = warp_sum_f32(qp);
                if (a.need_ortho) qo = warp_sum_f32(qo);
                if (my_c == c) { q32 = my_para ? qp : qo; trot_mine = trot; }
            }
            // ---- P2: per (component, spectrum) scalars in FP64, lanes <-> pairs ----
            if (pair_lane) {
                const NfSpecMeta &sm = a.spec[my_s];
                const double voff = ld_param<PT>(a.params, pbase + 0 * NC + my_c);
                double tex = ld_param<PT>(a.params, pbase + 2 * NC + my_c);
                const double ntot = ld_param<PT>(a.params, pbase + 3 * NC + my_c);
                const double sigm = ld_param<PT>(a.params, pbase + 4 * NC + my_c);
                const double orth = ld_param<PT>(a.params, pbase + 5 * NC + my_c);
                if (a.lte) tex = trot_mine;
                const double a1 = NF_HK * (2.0 * NF_BROT + (NF_CROT - NF_BROT));
                const double a2 = NF_HK * (6.0 * NF_BROT + 4.0 * (NF_CROT - NF_BROT));
                const double lev1 = 3.0 * fastexp_f64(a1 / trot_mine);
                const double lev2 = 5.0 * fastexp_f64(a2 / trot_mine);
                double zlev = my_J == 1 ? lev1 : lev2;
                if (my_J > 2) {
                    const double J = (double)my_J;
                    zlev = (2.0 * J + 1.0) *
                           fastexp_f64(NF_HK * (NF_BROT * J * (J + 1.0) + (NF_CROT - NF_BROT) * J * J) / trot_mine);
                }
                const double qtot = my_para ? lev1 + lev2 + (double)q32 : (double)q32;
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
            const double nu_min = sm.nu_min, inv_chan = sm.inv_chan;
            const float t0a = sm.t0a, t0b = sm.t0b;
            // ---- P3: per-line window + Gaussian coefficients, lanes <-> lines ----
            uint32_t lohi[NC];
            int keyF[NC], keyE[NC];     // NH3: first chunk at which a line has ended / has started
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int nl = IS_NH3 ? sm.nlines : ncomp;
                const bool act = lane < nl;
                double nucen, w;
                float wT;
                if (IS_NH3) {
                    const double f = act ? g_line_freq[sm.line_off + lane] : sm.nu0;
                    wT = act ? (float)sc.tauT[c][s] * (float)g_line_wt[sm.line_off + lane] : 0.0f;
                    w = sc.soc[c] * f;                  // hyperfine.pyx:71
                    nucen = f - sc.voc[c] * f;          // hyperfine.pyx:72-73
                } else {
                    const double f = sm.nu0;            // gaussian.pyx:28-33
                    const int cl = act ? lane : 0;
                    const double voff = ld_param<PT>(a.params, pbase + cl);
                    const double sigm = ld_param<PT>(a.params, pbase + ncomp + cl);
                    wT = act ? (float)ld_param<PT>(a.params, pbase + 2 * ncomp + cl) : 0.0f;
                    w = sigm / NF_CKMS * f;
                    nucen = f * (1.0 - voff / NF_CKMS);
                }
                const double cut = 5.0 * fabs(w);          // sqrt(12.5 / (0.5 / w^2)), hyperfine.pyx:82
                const double rel = nucen - nu_min;
                // floor((nu_cen - nu_min -/+ cut) / nu_chan), hyperfine.pyx:83-87.  cvt.rmi saturates
                // and maps NaN to 0, so non-finite parameters end up with an empty window.
                int lo = __double2int_rd((rel - cut) * inv_chan);
                int hi = __double2int_rd((rel + cut) * inv_chan);
                bool on = act && !(hi < 0 || lo > a.n_chan - 1);   // hyperfine.pyx:88
                const bool below = act && hi < 0;
                lo = max(lo, 0);
                hi = min(hi, a.n_chan - 1);
                const bool inband = on;
                on = on && hi > lo;                                // loop j in [lo, hi)
                if (IS_NH3) {
                    // keys are ascending in the (frequency-sorted) line index: windows entirely below
                    // the band have always ended, those above it (and padding lanes) never start
                    const int big = 1 << 24;
                    int E = below ? -big : (inband ? (lo >> 5) : big);
                    int F = below ? -big : (inband ? (on ? ((hi - 1) >> 5) + 1 : (lo >> 5)) : big);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {             // running max keeps F sorted around empty windows
                        const int v = __shfl_up_sync(NF_FULL, F, o);
                        if (lane >= o) F = max(F, v);
                    }
                    keyE[c] = E;
                    keyF[c] = max(F, E);
                }
                const double jc = rel * inv_chan;
                const int Ri = __double2int_rn(jc);
                const float phi = (float)(jc - (double)Ri);
                const float sch = (float)(w * inv_chan);
                const float k2 = __fdividef(0.5f * (float)NF_LOG2E, sch * sch);
                float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
                float2 Bw = make_float2(0.f, 0.f);
                if (on) {
                    A.x = (float)Ri;
                    A.y = -k2;
                    A.z = 2.0f * k2 * phi;
                    A.w = wT * ex2_approx(-k2 * phi * phi);
                    Bw.x = (float)(lo - Ri);
                    Bw.y = (float)(hi - Ri);
                }
                if constexpr (IS_NH3) {
                    // line i is element 0 of pair (p = i & 1, q = i >> 1) and element 1 of the pair
                    // of the other parity that starts one line earlier
                    float *e0 = reinterpret_cast<float *>(&sc.pair[c][lane & 1][lane >> 1]);
                    e0[0] = -A.x; e0[2] = A.y; e0[4] = A.z; e0[6] = A.w; e0[8] = Bw.x; e0[10] = Bw.y;
                    if (lane > 0) {
                        float *e1 = reinterpret_cast<float *>(&sc.pair[c][(lane - 1) & 1][(lane - 1) >> 1]);
                        e1[1] = -A.x; e1[3] = A.y; e1[5] = A.z; e1[7] = A.w; e1[9] = Bw.x; e1[11] = Bw.y;
                    } else {
                        // line 32 (element 1 of the last odd pair) and pair (32, 33) are null lines
                        float *z = reinterpret_cast<float *>(&sc.pair[c][1][15]);
                        z[1] = 0.f; z[3] = 0.f; z[5] = 0.f; z[7] = 0.f; z[9] = 0.f; z[11] = 0.f;
                        sc.pair[c][0][16].a = make_float4(0.f, 0.f, 0.f, 0.f);
                        sc.pair[c][0][16].b = make_float4(0.f, 0.f, 0.f, 0.f);
                        sc.pair[c][0][16].w = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                } else {
                    sc.line[c][lane].a = A;
                    sc.line[c][lane].w = Bw;
                }
                lohi[c] = on ? ((uint32_t)lo | ((uint32_t)hi << 16)) : 0u;
            }
            __syncwarp();
            if (!data_ready) { mbar_wait(bar, 0); data_ready = true; }

            float4 ampc[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) ampc[c] = IS_NH3 ? sc.amp[c][s] : make_float4(0.f, 0.f, 0.f, 0.f);

            // this pixel's row: the CTA's staged copy in shared memory when the vector belongs to
            // the tile's pixel, else straight from HBM/L2 (generic pointer, one LD per chunk)
            const float *drow = nullptr;
            if (have_data)
                drow = (pix == pix0 ? sdata + s * a.n_pad : a.data + pix * a.pix_stride + (int64_t)s * a.n_pad) + lane;
            float acc = 0.0f;
            if constexpr (IS_NH3) {
              for (int sb = 0; sb < nchunks; sb += 32) {
                // per-chunk dispatch table: lanes <-> chunks; the lines touching chunk g are the
                // contiguous run [#ended(g), #started(g)) of the frequency-sorted records
                {
                    uint32_t ent[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int g = sb + lane;
                        const int first = count_le_sorted(keyF[c], g);
                        const int end = count_le_sorted(keyE[c], g);
                        const int cnt = end - first;
                        const int poff = ((first & 1) * NF_MAX_PAIRS + (first >> 1)) * (int)sizeof(PairRec);
                        ent[c] = cnt > 0 ? ((uint32_t)poff | ((uint32_t)((cnt + 1) >> 1) << 16)) : 0u;
                    }
                    __syncwarp();
                    sc.tab[lane] = make_uint4(ent[0], ent[1], ent[2], ent[3]);
                    __syncwarp();
                }
                const int cend = min(32, nchunks - sb);
                float xj = (float)((sb << 5) + lane);
                const float *dp = have_data ? drow + (sb << 5) : nullptr;
                const uint4 *tp = sc.tab;
                for (int cc = 0; cc < cend; ++cc, xj += 32.0f, dp += 32, ++tp) {
                    const uint4 e4 = *tp;
                    const uint32_t ent[4] = {e4.x, e4.y, e4.z, e4.w};
                    float d = 0.0f;
                    if (have_data) d = *dp;
                    if ((e4.x | e4.y | e4.z | e4.w) == 0u) {   // no line of any component touches this chunk
                        if (WRITE_PRED) {
                            const int j = ((sb + cc) << 5) + lane;
                            if (j < a.n_chan) a.pred[(b * a.n_spec + s) * (int64_t)a.n_chan + j] = 0.0f;
                        }
                        acc = fmaf(d, d, acc);
                        continue;
                    }
                    const float T0 = fmaf(t0b, xj, t0a);
                    const uint64_t xj2 = pack2(xj, xj);
                    float m = 0.0f;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const uint32_t ec = ent[c];
                        if (ec == 0u) continue;
                        const PairRec *rec = reinterpret_cast<const PairRec *>(
                            reinterpret_cast<const unsigned char *>(&sc.pair[c][0][0]) + (ec & 0xffffu));
                        int n2 = (int)(ec >> 16);
                        float tau0 = 0.0f, tau1 = 0.0f;
                        // two lines per trip (packed FP32x2); a trailing odd slot holds the next line,
                        // whose own window test masks it off in this chunk
#pragma unroll 1
                        do {
                            pair_term(tau0, tau1, rec, xj2);
                            ++rec;
                        } while (--n2 > 0);
                        const float tau = tau0 + tau1;
                        const float4 am = ampc[c];
                        const float D = fmaxf(fmaf(am.y, xj, am.x), fmaf(am.w, xj, am.z));
                        // 1 - exp(-tau): FastExp's Taylor branch below 2^-5 (fastexp.c:265-270)
                        const float e1s = tau * (1.0f - 0.5f * tau * (1.0f - tau * (1.0f / 3.0f)));
                        const float e1l = 1.0f - ex2_approx(-(float)NF_LOG2E * tau);
                        const float e1 = fabsf(tau) < 0.03125f ? e1s : e1l;
                        m = fmaf(T0 * D, e1, m);
                    }
                    if (WRITE_PRED) {
                        const int j = ((sb + cc) << 5) + lane;
                        if (j < a.n_chan) a.pred[(b * a.n_spec + s) * (int64_t)a.n_chan + j] = m;
                    }
                    const float r = d - m;
                    acc = fmaf(r, r, acc);
                }
              }
            } else {
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
                uint32_t un_any = 0u;
#pragma unroll
                for (int c = 0; c < NC; ++c) un_any |= un[c];
                const int cend = min(32, nchunks - sb);
                float xj = (float)((sb << 5) + lane);
                for (int cc = 0; cc < cend; ++cc, xj += 32.0f) {
                    const int g = sb + cc;
                    float d = 0.0f;
                    if (have_data) d = drow[g << 5];
                    if (!((un_any >> cc) & 1u)) {          // no line of any component touches this chunk
                        if (WRITE_PRED) {
                            const int j = (g << 5) + lane;
                            if (j < a.n_chan) a.pred[(b * a.n_spec + s) * (int64_t)a.n_chan + j] = 0.0f;
                        }
                        acc = fmaf(d, d, acc);
                        continue;
                    }
                    float m = 0.0f;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        if (!((un[c] >> cc) & 1u)) continue;
                        uint32_t lm = __ballot_sync(NF_FULL, (cm[c] >> cc) & 1u);
                        float tau = 0.0f;
                        while (lm) {          // components are unordered: walk the set bits
                            const int i = __ffs(lm) - 1;
                            lm &= lm - 1;
                            line_term(tau, &sc.line[c][i], xj);
                        }
                        m += tau;
                    }
                    if (WRITE_PRED) {
                        const int j = (g << 5) + lane;
                        if (j < a.n_chan) a.pred[(b * a.n_spec + s) * (int64_t)a.n_chan + j] = m;
                    }
                    const float r = d - m;
                    acc = fmaf(r, r, acc);
                }
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

template <int NC, bool NH3>
static size_t like_smem_bytes(const NfLikeArgs &a)
{
    size_t data = (((size_t)a.n_spec * a.n_pad * 4 + 127) / 128) * 128;
    return 128 + data + sizeof(WarpScratch<NC, NH3>) * NF_WARPS_PER_CTA;
}

template <int NC, bool IS_NH3, bool WP, typename PT>
static cudaError_t launch_one(const NfLikeArgs &a, cudaStream_t st)
{
    auto kern = nf_like_kernel<NC, IS_NH3, WP, PT>;
    size_t smem = like_smem_bytes<NC, IS_NH3>(a);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    const int tile = a.tile_vecs > 0 ? a.tile_vecs : NF_TILE_VECS;
    int64_t grid = (a.B + tile - 1) / tile;
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

cudaError_t nf_launch_nh3_legacy(const NfLikeArgs &a, cudaStream_t st)
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

// sum of d^2 over every 32-channel chunk of every (pixel, spectrum) row: one warp per chunk
__global__ void nf_d2chunk_kernel(const float *data, float *out, int64_t n_chunks_total)
{
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_chunks_total) return;
    const float d = data[g * 32 + lane];
    const float v = warp_sum_f32(d * d);
    if (lane == 0) out[g] = v;
}

cudaError_t nf_launch_d2chunk(const float *data, float *out, int64_t n_rows, int n_pad, cudaStream_t st)
{
    const int64_t total = n_rows * (n_pad / 32);
    if (total <= 0) return cudaSuccess;
    const int wpb = 8;
    nf_d2chunk_kernel<<<(unsigned)((total + wpb - 1) / wpb), wpb * 32, 0, st>>>(data, out, total);
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
