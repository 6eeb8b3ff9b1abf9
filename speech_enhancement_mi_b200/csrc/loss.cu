// Loss terms of compute_loss (CRN_ELU.py:513-535; fullsubnet.py:964-987), forward and (se_loss_terms_grad) backward:
//   se_cal_si_snr  = utility.cal_si_snr (utility.py:207-223)
//   se_stoi_loss   = utility.stoi_loss (utility.py:821-916): torchaudio Resample 16k -> 10k (polyphase windowed sinc),
//                    removeSilentFrames (utility.py:521-571), Spectrogram(512, win 256, hop 128, power 2), 15 one-third
//                    octave bands (thirdoct, utility.py:480-518), 30-frame segments, clipping, correlation.
// One CTA per batch item walks the stages through a global workspace; these terms are ~1 % of a training step
// (SURVEY.md section 3.2), so the kernels favour exactness of the index work over speed.
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/se_b200.h"
#include "se_internal.h"

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

__global__ void __launch_bounds__(kThreads) stoi_kernel(const float* y_true, const float* y_pred, const int* lens,
                                                        long long L, StoiWork w, float* D) {
    __shared__ double red[kThreads / 32];
    __shared__ float redf[kThreads / 32];
    __shared__ float2 tw[512];
    __shared__ int s_ns;
    const int item = blockIdx.x, tid = threadIdx.x;
    const long long len = min((long long)lens[item], L);
    const float* src[2] = {y_true + (long long)item * L, y_pred + (long long)item * L};
    float* r10[2] = {w.r10 + ((long long)item * 2) * w.L10max, w.r10 + ((long long)item * 2 + 1) * w.L10max};
    float* sil[2] = {w.sil + ((long long)item * 2) * w.Lsmax, w.sil + ((long long)item * 2 + 1) * w.Lsmax};
    float* energy = w.energy + (long long)item * w.NFmax;
    int* sel = w.sel + (long long)item * w.NFmax;
    for (int i = tid; i < 512; i += blockDim.x) {
        float sn, cs;
        sincospif(-(float)i / 256.f, &sn, &cs);  // exp(-2 pi i k / 512)
        tw[i] = make_float2(cs, sn);
    }
    // (1) resample: out[5q + j] = sum_k kern[j][k] * wave[8q + k - 10], target length ceil(5 len / 8)
    const long long L10 = (5 * len + 7) / 8;
    for (int sgl = 0; sgl < 2; ++sgl)
        for (long long n = tid; n < L10; n += blockDim.x) {
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
    __syncthreads();
    // (2) removeSilentFrames: frames of 256 at hop 128 of the TRUE signal, energy, mask, compaction (order kept)
    const int n1 = (int)(L10 / 256), n2 = L10 >= 128 ? (int)((L10 - 128) / 256) : 0;
    const int NF = n1 + n2;
    if (NF == 0) {  // torch.max of an empty tensor raises -> the reference falls back to the raw signal, <= 512 samples
        if (tid == 0) D[item] = 0.99f;
        return;
    }
    float emax = -INFINITY;
    for (int m = tid; m < NF; m += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < 256; ++i) {
            const float v = c_st.hann_sym[i] * r10[0][128 * m + i];
            acc = fmaf(v, v, acc);
        }
        const float e = 20.f * log10f(sqrtf(acc) / 16.0f + (float)kEps64);
        energy[m] = e;
        emax = fmaxf(emax, e);
    }
    emax = block_max(emax, redf);
    if (tid == 0) {
        int ns = 0;
        for (int m = 0; m < NF; ++m)
            if (energy[m] - emax + 40.f > 0.f) sel[ns++] = m;
        s_ns = ns;
    }
    __syncthreads();
    const int ns = s_ns;
    const long long Ls = 128LL * (ns + 1);
    if (Ls <= 512) {
        if (tid == 0) D[item] = 0.99f;
        return;
    }
    for (int sgl = 0; sgl < 2; ++sgl)
        for (long long n = tid; n < Ls; n += blockDim.x) {
            const int k = (int)(n / 128), i = (int)(n % 128);
            float v = 0.f;
            if (k < ns) v += c_st.hann_sym[i] * r10[sgl][128 * sel[k] + i];                 // first half of frame k
            if (k >= 1) v += c_st.hann_sym[128 + i] * r10[sgl][128 * sel[k - 1] + 128 + i];  // second half of frame k-1
            sil[sgl][n] = v;
        }
    __syncthreads();
    // (3) power spectrogram (n_fft 512, hann(256) centred, hop 128, center=True reflect) -> one-third octave bands
    const int NS = 1 + (int)(Ls / 128);
    float* oct[2] = {w.oct + ((long long)item * 2) * 15 * w.NSmax, w.oct + ((long long)item * 2 + 1) * 15 * w.NSmax};
    const int klo = c_st.band_lo[0], khi = c_st.band_hi[14];
    for (int sgl = 0; sgl < 2; ++sgl) {
        for (int i = tid; i < 15 * NS; i += blockDim.x) oct[sgl][(i / NS) * w.NSmax + i % NS] = 0.f;
        __syncthreads();
        const int nb = khi - klo;
        for (int o = tid; o < NS * nb; o += blockDim.x) {
            const int r = o / nb, k = klo + o % nb;
            float re = 0.f, im = 0.f;
            for (int n = 0; n < 256; ++n) {  // the window is zero outside [128, 384) of the 512-sample frame
                long long pos = 128LL * r + 128 + n - 256;  // index into the un-padded signal
                if (pos < 0) pos = -pos;
                if (pos >= Ls) pos = 2 * (Ls - 1) - pos;
                const float v = sil[sgl][pos] * c_st.hann_per[n];
                const float2 t = tw[(k * (n + 128)) & 511];
                re = fmaf(v, t.x, re);
                im = fmaf(v, t.y, im);
            }
            const float pw = re * re + im * im;
            if (sgl == 1 && w.spec != nullptr) {
                float* sp = w.spec + (((long long)item * w.NSmax + r) * w.NBmax + (k - klo)) * 2;
                sp[0] = re;
                sp[1] = im;
            }
            for (int j = 0; j < 15; ++j)
                if (k >= c_st.band_lo[j] && k < c_st.band_hi[j]) atomicAdd(&oct[sgl][j * w.NSmax + r], pw);
        }
        __syncthreads();
        for (int i = tid; i < 15 * NS; i += blockDim.x) {
            float* p = &oct[sgl][(i / NS) * w.NSmax + i % NS];
            *p = sqrtf(*p + 1e-14f);
        }
        __syncthreads();
    }
    // (4) 30-frame segments: clip, normalise, correlate
    const int Nseg = 30;
    int M = NS - (Nseg - 1);
    int seglen = Nseg;
    if (M <= 0) {
        M = 1;
        seglen = NS;
    }
    const float c = 5.62341325f;
    double dsum = 0;
    for (int row = tid; row < 15 * M; row += blockDim.x) {
        const int m = row / 15, j = row % 15;
        const float* X = oct[0] + j * w.NSmax + m;
        const float* Y = oct[1] + j * w.NSmax + m;
        float nx = 0.f, ny = 0.f, mx = 0.f;
        for (int i = 0; i < seglen; ++i) {
            nx = fmaf(X[i], X[i], nx);
            ny = fmaf(Y[i], Y[i], ny);
            mx += X[i];
        }
        const float alpha = sqrtf(nx) / (sqrtf(ny) + (float)kEps64);
        mx /= seglen;
        float my = 0.f;
        for (int i = 0; i < seglen; ++i) my += fminf(Y[i] * alpha, X[i] + X[i] * c);
        my /= seglen;
        float sxx = 0.f, syy = 0.f, sxy = 0.f;
        for (int i = 0; i < seglen; ++i) {
            const float xc = X[i] - mx, yc = fminf(Y[i] * alpha, X[i] + X[i] * c) - my;
            sxx = fmaf(xc, xc, sxx);
            syy = fmaf(yc, yc, syy);
            sxy = fmaf(xc, yc, sxy);
        }
        dsum += sxy / ((sqrtf(sxx) + (float)kEps64) * (sqrtf(syy) + (float)kEps64));
    }
    const double tot = block_sum(dsum, red);
    if (tid == 0) D[item] = (float)(tot / (15.0 * M));
    if (w.dpred == nullptr) return;

    // ================= backward: d (-mean_B D) / d pred, stage by stage in reverse =====================================
    float* doct = w.doct + (long long)item * 15 * w.NSmax;
    float* dsil = w.dsil + (long long)item * w.Lsmax;
    float* dr10 = w.dr10 + (long long)item * w.L10max;
    int* rank = w.rank + (long long)item * w.NFmax;
    float* spec = w.spec + (long long)item * w.NSmax * w.NBmax * 2;
    for (int i = tid; i < 15 * w.NSmax; i += blockDim.x) doct[i] = 0.f;
    for (long long i = tid; i < Ls; i += blockDim.x) dsil[i] = 0.f;
    for (int m = tid; m < NF; m += blockDim.x) rank[m] = -1;
    __syncthreads();
    for (int k = tid; k < ns; k += blockDim.x) rank[sel[k]] = k;
    // (4') correlation rows -> d oct_pred
    const float wrow = w.gscale / (15.0f * M);
    for (int row = tid; row < 15 * M; row += blockDim.x) {
        const int m = row / 15, j = row % 15;
        const float* X = oct[0] + j * w.NSmax + m;
        const float* Y = oct[1] + j * w.NSmax + m;
        float nx = 0.f, ny = 0.f, mx = 0.f;
        for (int i = 0; i < seglen; ++i) {
            nx = fmaf(X[i], X[i], nx);
            ny = fmaf(Y[i], Y[i], ny);
            mx += X[i];
        }
        const float sny = sqrtf(ny), snx = sqrtf(nx);
        const float alpha = snx / (sny + (float)kEps64);
        mx /= seglen;
        float my = 0.f;
        for (int i = 0; i < seglen; ++i) my += fminf(Y[i] * alpha, X[i] + X[i] * c);
        my /= seglen;
        float sxx = 0.f, syy = 0.f, sxy = 0.f;
        for (int i = 0; i < seglen; ++i) {
            const float xc = X[i] - mx, yc = fminf(Y[i] * alpha, X[i] + X[i] * c) - my;
            sxx = fmaf(xc, xc, sxx);
            syy = fmaf(yc, yc, syy);
            sxy = fmaf(xc, yc, sxy);
        }
        const float ssy = sqrtf(syy);
        const float dx = sqrtf(sxx) + (float)kEps64, dy = ssy + (float)kEps64;
        const float c1 = 1.f / (dx * dy), c2 = ssy > 0.f ? sxy / (dx * dy * dy * ssy) : 0.f;
        // d value / d y_i = c1 xc_i - c2 yc_i (its mean over i vanishes because xc and yc are centred)
        float dalpha = 0.f;
        for (int i = 0; i < seglen; ++i) {
            const float ay = Y[i] * alpha, lim = X[i] + X[i] * c;
            if (ay < lim) dalpha += (c1 * (X[i] - mx) - c2 * (ay - my)) * Y[i];
        }
        const float da = sny > 0.f ? -dalpha * snx / ((sny + (float)kEps64) * (sny + (float)kEps64) * sny) : 0.f;
        for (int i = 0; i < seglen; ++i) {
            const float ay = Y[i] * alpha, lim = X[i] + X[i] * c;
            float g = da * Y[i];
            if (ay < lim) g += alpha * (c1 * (X[i] - mx) - c2 * (ay - my));
            atomicAdd(&doct[j * w.NSmax + m + i], wrow * g);
        }
    }
    __syncthreads();
    // (3') band energies -> power spectrum -> frames of the silent-frame-removed signal
    {
        const int nb = khi - klo;
        for (int o = tid; o < NS * nb; o += blockDim.x) {
            const int r = o / nb, k = klo + o % nb;
            float dP = 0.f;
            for (int j = 0; j < 15; ++j)
                if (k >= c_st.band_lo[j] && k < c_st.band_hi[j]) dP += doct[j * w.NSmax + r] / (2.f * oct[1][j * w.NSmax + r]);
            float* sp = spec + ((long long)r * w.NBmax + (k - klo)) * 2;
            sp[0] *= 2.f * dP;
            sp[1] *= 2.f * dP;
        }
        __syncthreads();
        for (int o = tid; o < NS * 256; o += blockDim.x) {
            const int r = o >> 8, n = o & 255;
            float dv = 0.f;
            const float* sp = spec + (long long)r * w.NBmax * 2;
            for (int kk = 0; kk < nb; ++kk) {
                const float2 t = tw[((klo + kk) * (n + 128)) & 511];
                dv = fmaf(sp[2 * kk], t.x, dv);
                dv = fmaf(sp[2 * kk + 1], t.y, dv);
            }
            long long pos = 128LL * r + 128 + n - 256;
            if (pos < 0) pos = -pos;
            if (pos >= Ls) pos = 2 * (Ls - 1) - pos;
            atomicAdd(&dsil[pos], dv * c_st.hann_per[n]);
        }
        __syncthreads();
    }
    // (2') overlap-add of the kept frames -> resampled signal
    for (long long p = tid; p < L10; p += blockDim.x) {
        const int m = (int)(p / 128), i = (int)(p % 128);
        float v = 0.f;
        if (m < NF && rank[m] >= 0) v += c_st.hann_sym[i] * dsil[128LL * rank[m] + i];
        if (m >= 1 && m - 1 < NF && rank[m - 1] >= 0) v += c_st.hann_sym[128 + i] * dsil[128LL * (rank[m - 1] + 1) + i];
        dr10[p] = v;
    }
    __syncthreads();
    // (1') polyphase resampler: out[5q + j] = sum_k kern[j][k] wave[8q + k - 10]
    float* dp = w.dpred + (long long)item * L;
    for (long long idx = tid; idx < len; idx += blockDim.x) {
        float acc = 0.f;
        for (long long q = (idx + 10) / 8; q >= 0; --q) {
            const long long k = idx + 10 - 8 * q;
            if (k >= 28) break;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const long long n = 5 * q + j;
                if (n < L10) acc = fmaf(c_st.resamp[j][k], dr10[n], acc);
            }
        }
        dp[idx] = acc;
    }
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
    SE_CUDA_OK(cudaMallocAsync(&base, sizeof(float) * (n_r10 + n_sil + n_en + n_oct + B) + sizeof(int) * n_en, st));
    w.r10 = base;
    w.sil = w.r10 + n_r10;
    w.energy = w.sil + n_sil;
    w.oct = w.energy + n_en;
    float* D = w.oct + n_oct;
    w.sel = reinterpret_cast<int*>(D + B);
    stoi_kernel<<<B, kThreads, 0, st>>>(y_true, y_pred, lens_dev, L, w, D);
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
    stoi_kernel<<<B, kThreads, 0, st>>>(source, pred, lens_dev, L, w, D);
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
