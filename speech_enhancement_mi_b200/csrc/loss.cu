// Loss terms of compute_loss (CRN_ELU.py:513-535; fullsubnet.py:964-987), forward and (se_loss_terms_grad) backward:
//   se_cal_si_snr  = utility.cal_si_snr (utility.py:207-223)
//   se_stoi_loss   = utility.stoi_loss (utility.py:821-916): torchaudio Resample 16k -> 10k (polyphase windowed sinc),
//                    removeSilentFrames (utility.py:521-571), Spectrogram(512, win 256, hop 128, power 2), 15 one-third
//                    octave bands (thirdoct, utility.py:480-518), 30-frame segments, clipping, correlation.
// One CTA per batch item walks the stages through a global workspace; these terms are ~1 % of a training step
// (SURVEY.md section 3.2), so the kernels favour exactness of the index work over speed.
#include <math.h>
#include <stdint.h>

#include <cooperative_groups.h>

#include <string>
#include <vector>

#include "../../include/se_b200.h"
#include <mutex>

#include "se_internal.h"

namespace cg = cooperative_groups;

namespace se {
namespace {

constexpr int kThreads = 512;
constexpr double kEps64 = 2.220446049250313e-16;  // np.finfo("float").eps (utility.py:477)

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = -INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = fmaxf(s, red[w]);
    return s;
}

// ---- SI-SNR --------------------------------------------------------------------------------------------------------
// out[i] = 20 log10(eps + |proj| / (|est_c - proj| + eps)), proj = <est_c, src_c> src_c / (|src_c|^2 + eps)
__global__ void __launch_bounds__(kThreads) si_snr_kernel(const float* est, const float* src, const int* len, long long L,
                                                          float eps, float* per_item, float* grad, float gscale) {
    __shared__ double red[kThreads / 32];
    const int i = blockIdx.x;
    const int n = len ? min((long long)len[i], L) : L;
    const float* e = est + (long long)i * L;
    const float* s = src + (long long)i * L;
    double se = 0, ss = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        se += e[k];
        ss += s[k];
    }
    const float me = (float)(block_sum(se, red) / n), ms = (float)(block_sum(ss, red) / n);
    double dot = 0, nss = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float ec = e[k] - me, sc = s[k] - ms;
        dot += (double)ec * sc;
        nss += (double)sc * sc;
    }
    const float fdot = (float)block_sum(dot, red);
    const float fnss = (float)block_sum(nss, red);
    const float alpha = fdot / (fnss + eps);  // l2norm(source)**2 + eps
    double np = 0, nr = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float ec = e[k] - me, sc = s[k] - ms;
        const float tr = alpha * sc;
        np += (double)tr * tr;
        nr += (double)(ec - tr) * (ec - tr);
    }
    const float fp = sqrtf((float)block_sum(np, red)), fr = sqrtf((float)block_sum(nr, red));
    if (threadIdx.x == 0) per_item[i] = 20.f * log10f(eps + fp / (fr + eps));
    if (grad == nullptr) return;
    // backward: val = 20 log10(eps + q), q = fp / (fr + eps), tr = alpha sc, r = ec - tr, alpha = <ec, sc> / (nss + eps)
    const float q = fp / (fr + eps);
    const float dq = 20.f / (2.302585093f * (eps + q)) * gscale;
    const float dfp = dq / (fr + eps), dfr = -dq * fp / ((fr + eps) * (fr + eps));
    const float a_r = fr > 0.f ? dfr / fr : 0.f;                      // d val / d r[k] = a_r * r[k]
    const float tr_sc = alpha * fnss, r_sc = fdot - alpha * fnss;     // <tr, sc>, <r, sc>
    const float dalpha = (fp > 0.f ? dfp * tr_sc / fp : 0.f) - a_r * r_sc;
    const float a_s = dalpha / (fnss + eps);                          // + a_s * sc[k]
    double gs = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float ec = e[k] - me, sc = s[k] - ms;
        gs += (double)(a_r * (ec - alpha * sc) + a_s * sc);
    }
    const float gmean = (float)(block_sum(gs, red) / n);             // ec = e - mean(e)
    float* g = grad + (long long)i * L;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float ec = e[k] - me, sc = s[k] - ms;
        g[k] = a_r * (ec - alpha * sc) + a_s * sc - gmean;
    }
}

__global__ void mean_kernel(const float* v, int n, float scale, float* out) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += v[i];
    *out = scale * s / n;
}

// ---- STOI ----------------------------------------------------------------------------------------------------------
struct StoiTables {
    float resamp[5][28];  // torchaudio sinc_interp_hann kernel for 16000 -> 10000 (orig 8, new 5, width 10)
    float hann_sym[256];  // np.hanning(256)
    float hann_per[256];  // torch.hann_window(256) (periodic), centred in the 512-point frame
    int band_lo[15], band_hi[15];
};
__constant__ StoiTables c_st;

struct StoiWork {
    float* r10;     // [B][2][L10max] resampled true / pred
    float* sil;     // [B][2][Lsmax] silent-frame-removed signals
    float* energy;  // [B][NFmax]
    int* sel;       // [B][NFmax]
    float* oct;     // [B][2][15][NSmax]
    long long L10max, Lsmax;
    int NFmax, NSmax;
    // backward (null: forward only)
    float* spec;    // [B][NSmax][NBmax][2] spectrum of the prediction, then 2 dP re / 2 dP im
    float* doct;    // [B][15][NSmax]
    float* dsil;    // [B][Lsmax]
    float* dr10;    // [B][L10max]
    int* rank;      // [B][NFmax] position of frame m among the kept frames, -1 = silent
    float* dpred;   // [B][L] d (-mean D) / d pred
    int NBmax;
    float gscale;   // -1 / B
};

// One cooperative grid of (G, B) CTAs: the G CTAs of an item share its heavy loops (resampler, spectrogram DFTs and
// their adjoints) and separate the stages with grid barriers; the light serial stages (frame energies, silent-frame
// selection) are recomputed by every CTA in shared memory.  Items that the reference scores 0.99 (too short) stay in
// the grid (they must reach every barrier) but do no work.
__global__ void __launch_bounds__(kThreads) stoi_kernel(const float* y_true, const float* y_pred, const int* lens,
                                                        long long L, StoiWork w, float* D, double* dacc) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smraw[];
    float* s_energy = reinterpret_cast<float*>(smraw);       // [NFmax]
    int* s_sel = reinterpret_cast<int*>(s_energy + w.NFmax);  // [NFmax]
    int* s_rank = s_sel + w.NFmax;                            // [NFmax]
    float* s_v = reinterpret_cast<float*>(s_rank + w.NFmax);  // [256] windowed frame / its gradient
    float* s_pw = s_v + 256;                                  // [2*257] power per bin, then (2 dP re, 2 dP im)
    __shared__ double red[kThreads / 32];
    __shared__ float redf[kThreads / 32];
    __shared__ float2 tw[512];
    __shared__ int s_ns;
    const int item = blockIdx.y, tid = threadIdx.x, G = gridDim.x, g = blockIdx.x;
    const long long it0 = (long long)g * blockDim.x + tid, stride = (long long)G * blockDim.x;
    const long long len = min((long long)lens[item], L);
    const float* src[2] = {y_true + (long long)item * L, y_pred + (long long)item * L};
    float* r10[2] = {w.r10 + ((long long)item * 2) * w.L10max, w.r10 + ((long long)item * 2 + 1) * w.L10max};
    float* sil[2] = {w.sil + ((long long)item * 2) * w.Lsmax, w.sil + ((long long)item * 2 + 1) * w.Lsmax};
    float* oct[2] = {w.oct + ((long long)item * 2) * 15 * w.NSmax, w.oct + ((long long)item * 2 + 1) * 15 * w.NSmax};
    const bool want_grad = w.dpred != nullptr;
    float* doct = want_grad ? w.doct + (long long)item * 15 * w.NSmax : nullptr;
    float* dsil = want_grad ? w.dsil + (long long)item * w.Lsmax : nullptr;
    float* dr10 = want_grad ? w.dr10 + (long long)item * w.L10max : nullptr;
    float* spec = want_grad ? w.spec + (long long)item * w.NSmax * w.NBmax * 2 : nullptr;
    for (int i = tid; i < 512; i += blockDim.x) {
        float sn, cs;
        sincospif(-(float)i / 256.f, &sn, &cs);  // exp(-2 pi i k / 512)
        tw[i] = make_float2(cs, sn);
    }
    // (1) resample: out[5q + j] = sum_k kern[j][k] * wave[8q + k - 10], target length ceil(5 len / 8)
    const long long L10 = (5 * len + 7) / 8;
    for (int sgl = 0; sgl < 2; ++sgl)
        for (long long n = it0; n < L10; n += stride) {
            const long long q = n / 5;
            const int j = (int)(n - 5 * q);
            float acc = 0.f;
#pragma unroll 4
            for (int k = 0; k < 28; ++k) {
                const long long idx = 8 * q + k - 10;
                if (idx >= 0 && idx < len) acc = fmaf(c_st.resamp[j][k], src[sgl][idx], acc);
            }
            r10[sgl][n] = acc;
        }
    if (want_grad) {
        for (long long i = it0; i < 15LL * w.NSmax; i += stride) doct[i] = 0.f;
        for (long long i = it0; i < w.Lsmax; i += stride) dsil[i] = 0.f;
    }
    grid.sync();
    // (2) removeSilentFrames: frames of 256 at hop 128 of the TRUE signal, energy, mask, compaction (order kept)
    const int n1 = (int)(L10 / 256), n2 = L10 >= 128 ? (int)((L10 - 128) / 256) : 0;
    const int NF = n1 + n2;
    bool alive = NF > 0;  // torch.max of an empty tensor raises -> the reference falls back to the raw signal, <= 512 samples
    int ns = 0;
    long long Ls = 0;
    if (alive) {
        float emax = -INFINITY;
        for (int m = tid; m < NF; m += blockDim.x) {
            float acc = 0.f;
            for (int i = 0; i < 256; ++i) {
                const float v = c_st.hann_sym[i] * __ldcg(&r10[0][128 * m + i]);
                acc = fmaf(v, v, acc);
            }
            const float e = 20.f * log10f(sqrtf(acc) / 16.0f + (float)kEps64);
            s_energy[m] = e;
            s_rank[m] = -1;
            emax = fmaxf(emax, e);
        }
        emax = block_max(emax, redf);
        if (tid == 0) {
            int k = 0;
            for (int m = 0; m < NF; ++m)
                if (s_energy[m] - emax + 40.f > 0.f) {
                    s_rank[m] = k;
                    s_sel[k++] = m;
                }
            s_ns = k;
        }
        __syncthreads();
        ns = s_ns;
        Ls = 128LL * (ns + 1);
        alive = Ls > 512;
    }
    if (alive) {
        for (int sgl = 0; sgl < 2; ++sgl)
            for (long long n = it0; n < Ls; n += stride) {
                const int k = (int)(n / 128), i = (int)(n % 128);
                float v = 0.f;
                if (k < ns) v += c_st.hann_sym[i] * __ldcg(&r10[sgl][128 * s_sel[k] + i]);                 // first half of frame k
                if (k >= 1) v += c_st.hann_sym[128 + i] * __ldcg(&r10[sgl][128 * s_sel[k - 1] + 128 + i]);  // second half of frame k-1
                sil[sgl][n] = v;
            }
    }
    grid.sync();
    // (3) power spectrogram (n_fft 512, hann(256) centred, hop 128, center=True reflect) -> one-third octave bands;
    //     one CTA per frame, two threads per bin (each sums half of the 256 window samples)
    const int NS = alive ? 1 + (int)(Ls / 128) : 0;
    const int klo = c_st.band_lo[0], khi = c_st.band_hi[14];
    const int nb = khi - klo;
    for (int sgl = 0; sgl < 2 && alive; ++sgl) {
        for (int r = g; r < NS; r += G) {
            if (tid < 256) {
                long long pos = 128LL * r + 128 + tid - 256;  // index into the un-padded signal
                if (pos < 0) pos = -pos;
                if (pos >= Ls) pos = 2 * (Ls - 1) - pos;
                s_v[tid] = __ldcg(&sil[sgl][pos]) * c_st.hann_per[tid];
            }
            __syncthreads();
            {
                const int kk = tid >> 1, h = tid & 1;
                float re = 0.f, im = 0.f;
                if (kk < nb) {
                    const int k = klo + kk;
                    for (int n = 128 * h; n < 128 * h + 128; ++n) {  // the window is zero outside [128, 384) of the frame
                        const float2 t = tw[(k * (n + 128)) & 511];
                        re = fmaf(s_v[n], t.x, re);
                        im = fmaf(s_v[n], t.y, im);
                    }
                }
                re += __shfl_xor_sync(0xffffffffu, re, 1);
                im += __shfl_xor_sync(0xffffffffu, im, 1);
                if (kk < nb && h == 0) {
                    s_pw[kk] = re * re + im * im;
                    if (sgl == 1 && want_grad) {
                        float* sp = spec + ((long long)r * w.NBmax + kk) * 2;
                        sp[0] = re;
                        sp[1] = im;
                    }
                }
            }
            __syncthreads();
            if (tid < 15) {
                float acc = 0.f;
                for (int k = c_st.band_lo[tid]; k < c_st.band_hi[tid]; ++k) acc += s_pw[k - klo];
                oct[sgl][tid * w.NSmax + r] = sqrtf(acc + 1e-14f);
            }
            __syncthreads();
        }
    }
    grid.sync();
    // (4) 30-frame segments: clip, normalise, correlate (and, for the backward, d oct_pred)
    const int Nseg = 30;
    int M = NS - (Nseg - 1);
    int seglen = Nseg;
    if (M <= 0) {
        M = 1;
        seglen = NS;
    }
    const float c = 5.62341325f;
    double dsum = 0;
    const float wrow = w.gscale / (15.0f * M);
    for (long long row = it0; row < 15LL * M && alive; row += stride) {
        const int m = (int)(row / 15), j = (int)(row % 15);
        const float* X = oct[0] + j * w.NSmax + m;
        const float* Y = oct[1] + j * w.NSmax + m;
        float nx = 0.f, ny = 0.f, mx = 0.f;
        for (int i = 0; i < seglen; ++i) {
            const float x = __ldcg(X + i), y = __ldcg(Y + i);
            nx = fmaf(x, x, nx);
            ny = fmaf(y, y, ny);
            mx += x;
        }
        const float sny = sqrtf(ny), snx = sqrtf(nx);
        const float alpha = snx / (sny + (float)kEps64);
        mx /= seglen;
        float my = 0.f;
        for (int i = 0; i < seglen; ++i) {
            const float x = __ldcg(X + i);
            my += fminf(__ldcg(Y + i) * alpha, x + x * c);
        }
        my /= seglen;
        float sxx = 0.f, syy = 0.f, sxy = 0.f;
        for (int i = 0; i < seglen; ++i) {
            const float x = __ldcg(X + i);
            const float xc = x - mx, yc = fminf(__ldcg(Y + i) * alpha, x + x * c) - my;
            sxx = fmaf(xc, xc, sxx);
            syy = fmaf(yc, yc, syy);
            sxy = fmaf(xc, yc, sxy);
        }
        const float ssy = sqrtf(syy);
        const float dx = sqrtf(sxx) + (float)kEps64, dy = ssy + (float)kEps64;
        dsum += sxy / (dx * dy);
        if (want_grad) {
            const float c1 = 1.f / (dx * dy), c2 = ssy > 0.f ? sxy / (dx * dy * dy * ssy) : 0.f;
            // d value / d y_i = c1 xc_i - c2 yc_i (its mean over i vanishes because xc and yc are centred)
            float dalpha = 0.f;
            for (int i = 0; i < seglen; ++i) {
                const float x = __ldcg(X + i), y = __ldcg(Y + i);
                const float ay = y * alpha, lim = x + x * c;
                if (ay < lim) dalpha += (c1 * (x - mx) - c2 * (ay - my)) * y;
            }
            const float da = sny > 0.f ? -dalpha * snx / ((sny + (float)kEps64) * (sny + (float)kEps64) * sny) : 0.f;
            for (int i = 0; i < seglen; ++i) {
                const float x = __ldcg(X + i), y = __ldcg(Y + i);
                const float ay = y * alpha, lim = x + x * c;
                float gq = da * y;
                if (ay < lim) gq += alpha * (c1 * (x - mx) - c2 * (ay - my));
                atomicAdd(&doct[j * w.NSmax + m + i], wrow * gq);
            }
        }
    }
    const double tot = block_sum(dsum, red);
    if (tid == 0 && alive) atomicAdd(dacc + item, tot);
    grid.sync();
    if (g == 0 && tid == 0) D[item] = alive ? (float)(__ldcg(dacc + item) / (15.0 * M)) : 0.99f;
    if (!want_grad) return;

    // ================= backward: d (-mean_B D) / d pred, stage by stage in reverse =====================================
    // (3') band energies -> power spectrum -> frames of the silent-frame-removed signal
    for (int r = g; r < NS && alive; r += G) {
        if (tid < nb) {
            const int k = klo + tid;
            float dP = 0.f;
            for (int j = 0; j < 15; ++j)
                if (k >= c_st.band_lo[j] && k < c_st.band_hi[j])
                    dP += __ldcg(&doct[j * w.NSmax + r]) / (2.f * __ldcg(&oct[1][j * w.NSmax + r]));
            const float* sp = spec + ((long long)r * w.NBmax + tid) * 2;
            s_pw[2 * tid] = 2.f * dP * sp[0];
            s_pw[2 * tid + 1] = 2.f * dP * sp[1];
        }
        __syncthreads();
        {
            const int n = tid >> 1, h = tid & 1;  // two threads per window sample, each sums half of the bins
            const int k0 = h ? nb / 2 : 0, k1 = h ? nb : nb / 2;
            float dv = 0.f;
            for (int kk = k0; kk < k1; ++kk) {
                const float2 t = tw[((klo + kk) * (n + 128)) & 511];
                dv = fmaf(s_pw[2 * kk], t.x, dv);
                dv = fmaf(s_pw[2 * kk + 1], t.y, dv);
            }
            dv += __shfl_xor_sync(0xffffffffu, dv, 1);
            if (h == 0) {
                long long pos = 128LL * r + 128 + n - 256;
                if (pos < 0) pos = -pos;
                if (pos >= Ls) pos = 2 * (Ls - 1) - pos;
                atomicAdd(&dsil[pos], dv * c_st.hann_per[n]);
            }
        }
        __syncthreads();
    }
    grid.sync();
    // (2') overlap-add of the kept frames -> resampled signal
    for (long long p = it0; p < L10 && alive; p += stride) {
        const int m = (int)(p / 128), i = (int)(p % 128);
        float v = 0.f;
        if (m < NF && s_rank[m] >= 0) v += c_st.hann_sym[i] * __ldcg(&dsil[128LL * s_rank[m] + i]);
        if (m >= 1 && m - 1 < NF && s_rank[m - 1] >= 0)
            v += c_st.hann_sym[128 + i] * __ldcg(&dsil[128LL * (s_rank[m - 1] + 1) + i]);
        dr10[p] = v;
    }
    grid.sync();
    // (1') polyphase resampler: out[5q + j] = sum_k kern[j][k] wave[8q + k - 10]
    float* dp = w.dpred + (long long)item * L;
    for (long long idx = it0; idx < len && alive; idx += stride) {
        float acc = 0.f;
        for (long long q = (idx + 10) / 8; q >= 0; --q) {
            const long long k = idx + 10 - 8 * q;
            if (k >= 28) break;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const long long n = 5 * q + j;
                if (n < L10) acc = fmaf(c_st.resamp[j][k], __ldcg(&dr10[n]), acc);
            }
        }
        dp[idx] = acc;
    }
}

// cooperative launch: G CTAs per item, all co-resident
int launch_stoi(const float* y_true, const float* y_pred, const int* lens, int B, long long L, const StoiWork& w, float* D,
                double* dacc, cudaStream_t st) {
    const size_t smem = (size_t)w.NFmax * (sizeof(float) + 2 * sizeof(int)) + (256 + 2 * 257) * sizeof(float);
    int dev = 0, sms = 0, per_sm = 0;
    SE_CUDA_OK(cudaGetDevice(&dev));
    SE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stoi_kernel, kThreads, smem));
    SE_REQUIRE(per_sm >= 1, "stoi: kernel does not fit on an SM");
    int G = sms * per_sm / B;
    if (G > 96) G = 96;
    SE_REQUIRE(G >= 1, "stoi: batch larger than one co-resident grid (" + std::to_string(sms * per_sm) + " items)");
    SE_CUDA_OK(cudaMemsetAsync(dacc, 0, sizeof(double) * B, st));
    const float* a0 = y_true;
    const float* a1 = y_pred;
    const int* a2 = lens;
    long long a3 = L;
    StoiWork a4 = w;
    float* a5 = D;
    double* a6 = dacc;
    void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6};
    SE_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(stoi_kernel), dim3(G, B), dim3(kThreads), args, smem, st));
    return 0;
}

bool g_tables_ready[64] = {};

int upload_tables() {
    int dev = 0;
    SE_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 64 && g_tables_ready[dev]) return 0;
    StoiTables t{};
    const double pi = 3.14159265358979323846;
    // torchaudio.functional._get_sinc_resample_kernel(16000, 10000, gcd 2000, lowpass_filter_width=6, rolloff=0.99, hann)
    const int orig = 8, nw = 5, lpw = 6;
    const double base = nw * 0.99;
    const int width = (int)ceil(lpw * orig / base);  // 10
    for (int j = 0; j < nw; ++j)
        for (int k = 0; k < 2 * width + orig; ++k) {
            double tt = (-(double)j / nw + (double)(k - width) / orig) * base;
            if (tt < -lpw) tt = -lpw;
            if (tt > lpw) tt = lpw;
            const double win = cos(tt * pi / lpw / 2) * cos(tt * pi / lpw / 2);
            const double x = tt * pi;
            const double snc = x == 0.0 ? 1.0 : sin(x) / x;
            t.resamp[j][k] = (float)(snc * win * (base / orig));
        }
    for (int i = 0; i < 256; ++i) {
        t.hann_sym[i] = (float)(0.5 - 0.5 * cos(2 * pi * i / 255.0));  // np.hanning
        t.hann_per[i] = (float)(0.5 - 0.5 * cos(2 * pi * i / 256.0));  // torch.hann_window (periodic)
    }
    // thirdoct(fs=10000, nfft=512, num_bands=15, min_freq=150): f = linspace(0, fs, nfft+1)[:nfft/2+1] in float32
    std::vector<float> f(257);
    for (int i = 0; i < 257; ++i) f[i] = (float)(10000.0 * i / 512.0);
    for (int b = 0; b < 15; ++b) {
        const double lo = 150.0 * pow(2.0, (2.0 * b - 1) / 6.0), hi = 150.0 * pow(2.0, (2.0 * b + 1) / 6.0);
        int il = 0, ih = 0;
        double bl = 1e300, bh = 1e300;
        for (int i = 0; i < 257; ++i) {  // argmin of the squared distance, first minimum
            const double dl = ((double)f[i] - lo) * ((double)f[i] - lo), dh = ((double)f[i] - hi) * ((double)f[i] - hi);
            if (dl < bl) {
                bl = dl;
                il = i;
            }
            if (dh < bh) {
                bh = dh;
                ih = i;
            }
        }
        t.band_lo[b] = il;
        t.band_hi[b] = ih;
    }
    SE_CUDA_OK(cudaMemcpyToSymbol(c_st, &t, sizeof(t)));
    if (dev < 64) g_tables_ready[dev] = true;
    return 0;
}

// The loss entry points take their scratch from the stream-ordered allocator.  The device's default pool releases unused
// memory back to the OS at every synchronisation (release threshold 0), so a training loop that reads the loss on the host
// after each micro-step (train.py:199-206) paid a fresh physical allocation -- milliseconds -- in the next step.  Keep the
// pool's memory: once per device.
int keep_pool_memory() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    SE_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && !done[dev]) {
        cudaMemPool_t pool;
        SE_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        SE_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        done[dev] = true;
    }
    return 0;
}

}  // namespace
}  // namespace se

using namespace se;

extern "C" {

int se_cal_si_snr(const float* separated, const float* source, const int32_t* length_dev, int B, int64_t L, float* out,
                  void* stream) {
    SE_REQUIRE(separated && source && out, "se_cal_si_snr: null buffer");
    SE_REQUIRE(B > 0 && L > 0, "se_cal_si_snr: empty batch");
    cudaStream_t st = (cudaStream_t)stream;
    float* per = nullptr;
    if (keep_pool_memory()) return 1;
    SE_CUDA_OK(cudaMallocAsync(&per, sizeof(float) * B, st));
    si_snr_kernel<<<B, kThreads, 0, st>>>(separated, source, length_dev, L, 1e-8f, per, nullptr, 0.f);
    mean_kernel<<<1, 1, 0, st>>>(per, B, 1.0f, out);
    SE_CUDA_OK(cudaGetLastError());
    SE_CUDA_OK(cudaFreeAsync(per, st));
    return 0;
}

int se_stoi_loss(const float* y_true, const float* y_pred, const int32_t* lens_dev, int B, int64_t L, float* out,
                 void* stream) {
    SE_REQUIRE(y_true && y_pred && lens_dev && out, "se_stoi_loss: null buffer");
    SE_REQUIRE(B > 0 && L > 0, "se_stoi_loss: empty batch");
    if (upload_tables()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    StoiWork w{};
    w.L10max = (5 * L + 7) / 8 + 8;
    w.NFmax = (int)(w.L10max / 128) + 2;
    w.Lsmax = 128LL * (w.NFmax + 2);
    w.NSmax = (int)(w.Lsmax / 128) + 2;
    float* base = nullptr;
    const size_t n_r10 = (size_t)B * 2 * w.L10max, n_sil = (size_t)B * 2 * w.Lsmax, n_en = (size_t)B * w.NFmax,
                 n_oct = (size_t)B * 2 * 15 * w.NSmax;
    if (keep_pool_memory()) return 1;
    SE_CUDA_OK(cudaMallocAsync(&base, sizeof(float) * (n_r10 + n_sil + n_en + n_oct + B) + sizeof(int) * n_en, st));
    w.r10 = base;
    w.sil = w.r10 + n_r10;
    w.energy = w.sil + n_sil;
    w.oct = w.energy + n_en;
    float* D = w.oct + n_oct;
    w.sel = reinterpret_cast<int*>(D + B);
    double* dacc = nullptr;
    SE_CUDA_OK(cudaMallocAsync(&dacc, sizeof(double) * B, st));
    if (launch_stoi(y_true, y_pred, lens_dev, B, L, w, D, dacc, st)) return 1;
    SE_CUDA_OK(cudaFreeAsync(dacc, st));
    mean_kernel<<<1, 1, 0, st>>>(D, B, -1.0f, out);  // reduction="mean": -D.mean()
    SE_CUDA_OK(cudaGetLastError());
    SE_CUDA_OK(cudaFreeAsync(base, st));
    return 0;
}

int se_loss_terms_grad(const float* source, const float* pred, const int32_t* lens_dev, int B, int64_t L, float* out2,
                       float* d_stoi, float* d_sisnr, void* stream) {
    SE_REQUIRE(source && pred && lens_dev && out2 && d_stoi && d_sisnr, "se_loss_terms_grad: null buffer");
    SE_REQUIRE(B > 0 && L > 0, "se_loss_terms_grad: empty batch");
    if (upload_tables()) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    SE_CUDA_OK(cudaMemsetAsync(d_stoi, 0, sizeof(float) * (size_t)B * L, st));
    SE_CUDA_OK(cudaMemsetAsync(d_sisnr, 0, sizeof(float) * (size_t)B * L, st));
    StoiWork w{};
    w.L10max = (5 * L + 7) / 8 + 8;
    w.NFmax = (int)(w.L10max / 128) + 2;
    w.Lsmax = 128LL * (w.NFmax + 2);
    w.NSmax = (int)(w.Lsmax / 128) + 2;
    w.NBmax = 257;
    float* base = nullptr;
    const size_t n_r10 = (size_t)B * 2 * w.L10max, n_sil = (size_t)B * 2 * w.Lsmax, n_en = (size_t)B * w.NFmax,
                 n_oct = (size_t)B * 2 * 15 * w.NSmax, n_spec = (size_t)B * w.NSmax * w.NBmax * 2,
                 n_doct = (size_t)B * 15 * w.NSmax, n_dsil = (size_t)B * w.Lsmax, n_dr10 = (size_t)B * w.L10max;
    if (keep_pool_memory()) return 1;
    SE_CUDA_OK(cudaMallocAsync(&base,
                               sizeof(float) * (n_r10 + n_sil + n_en + n_oct + n_spec + n_doct + n_dsil + n_dr10 + 2 * B) +
                                   sizeof(int) * 2 * n_en,
                               st));
    w.r10 = base;
    w.sil = w.r10 + n_r10;
    w.energy = w.sil + n_sil;
    w.oct = w.energy + n_en;
    w.spec = w.oct + n_oct;
    w.doct = w.spec + n_spec;
    w.dsil = w.doct + n_doct;
    w.dr10 = w.dsil + n_dsil;
    float* D = w.dr10 + n_dr10;
    float* per = D + B;
    w.sel = reinterpret_cast<int*>(per + B);
    w.rank = w.sel + n_en;
    w.dpred = d_stoi;
    w.gscale = -1.0f / B;
    double* dacc = nullptr;
    SE_CUDA_OK(cudaMallocAsync(&dacc, sizeof(double) * B, st));
    if (launch_stoi(source, pred, lens_dev, B, L, w, D, dacc, st)) return 1;
    SE_CUDA_OK(cudaFreeAsync(dacc, st));
    mean_kernel<<<1, 1, 0, st>>>(D, B, -1.0f, out2);
    si_snr_kernel<<<B, kThreads, 0, st>>>(pred, source, lens_dev, L, 1e-8f, per, d_sisnr, 1.0f / B);
    mean_kernel<<<1, 1, 0, st>>>(per, B, 1.0f, out2 + 1);
    SE_CUDA_OK(cudaGetLastError());
    SE_CUDA_OK(cudaFreeAsync(base, st));
    return 0;
}

static __global__ void axpby_kernel(const float* a, const float* x, const float* b, const float* y, float* out, long long n) {
    const float fa = *a, fb = *b;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fa * x[i] + fb * y[i];
}
static __global__ void __launch_bounds__(512) sqnorm_kernel(const float* g, long long n, double* acc) {
    __shared__ double red[16];
    double s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += (double)g[i] * g[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(acc, s);
}
// torch.nn.utils.clip_grad_norm_ (coef = max_norm / (norm + 1e-6), clamped to 1) + torch.optim.Adam (no weight decay)
static __global__ void clip_adam_kernel(float* theta, float* grad, float* m, float* v, long long n, float lr, float b1, float b2,
                                 float eps, float bc1, float bc2, float max_norm, float gscale, const double* sq,
                                 float* norm_out) {
    const float norm = (float)sqrt(*sq) * gscale;
    float coef = gscale;
    if (max_norm > 0.f) coef *= fminf(1.f, max_norm / (norm + 1e-6f));
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float g = grad[i] * coef;
        grad[i] = g;
        const float mi = b1 * m[i] + (1.f - b1) * g;
        const float vi = b2 * v[i] + (1.f - b2) * g * g;
        m[i] = mi;
        v[i] = vi;
        theta[i] -= lr / bc1 * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    }
}

int se_axpby_dev(const float* a, const float* x, const float* b, const float* y, float* out, int64_t n, void* stream) {
    SE_REQUIRE(a && x && b && y && out, "se_axpby_dev: null buffer");
    if (n <= 0) return 0;
    long long g = (n + 255) / 256;
    axpby_kernel<<<(int)(g > 1184 ? 1184 : g), 256, 0, (cudaStream_t)stream>>>(a, x, b, y, out, n);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int se_clip_adam_step(float* theta, float* grad, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                      float eps, int step, float max_norm, float grad_scale, float* norm_out, void* stream) {
    SE_REQUIRE(theta && grad && m && v, "se_clip_adam_step: null buffer");
    SE_REQUIRE(n > 0 && step >= 1, "se_clip_adam_step: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    double* sq = nullptr;
    if (keep_pool_memory()) return 1;
    SE_CUDA_OK(cudaMallocAsync(&sq, sizeof(double), st));
    SE_CUDA_OK(cudaMemsetAsync(sq, 0, sizeof(double), st));
    long long g = (n + 511) / 512;
    sqnorm_kernel<<<(int)(g > 592 ? 592 : g), 512, 0, st>>>(grad, n, sq);
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    g = (n + 255) / 256;
    clip_adam_kernel<<<(int)(g > 1184 ? 1184 : g), 256, 0, st>>>(theta, grad, m, v, n, lr, beta1, beta2, eps, bc1, bc2, max_norm,
                                                                 grad_scale, sq, norm_out);
    SE_CUDA_OK(cudaGetLastError());
    SE_CUDA_OK(cudaFreeAsync(sq, st));
    return 0;
}

}  // extern "C"
