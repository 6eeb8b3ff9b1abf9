// One transposed-conv decoder block of the CRN (CRN_ELU.py:290-307) for the levels with few channels, fp16 operand mode,
// one stream per CTA at a time with the stream's tensors resident in shared memory:
//
//     y  = ELU(ConvTranspose2d(x)[..., -T:])           both output parities from one pass over the input rows
//     m  = sigmoid(GLN_r(W_m s + b_m)),  r = ELU(W_r s + b_r)          s = skip tensor (an encoder block's output)
//     out = m * r + (1 - m) * GLN(y)                   (bins >= 2 Fin - 1: GLN(y) := 0, the zero padding of :299-303)
//
// Before: three launches (merged-parity tcgen05 GEMM -> y; 1x1 pair kernel -> mask, residual; blend kernel) with y, the
// mask pre-activation and the residual making a round trip through HBM (0.24 ms for 32 -> 16 channels, 1024 streams).
// Here y stays in shared memory, the skip pair is recomputed instead of stored (K = 16: one mma per tile and kind), both
// GlobalLayerNorm statistics are reduced inside the CTA and only the block output is written.
//
// Transposed conv as a GEMM over the INPUT rows (t, f'): output bin 2 f' uses frequency taps 0, 2, 4 and bin 2 f' + 1
// taps 1, 3, all on the padded input rows f' + 2 - j (j = 0..2); the `[..., -T:]` crop turns the time taps into
// look-ahead, frame t + (2 - kt) d.  K = 9 (kt, j) taps x Cin, N = [even | odd] x Cout; the odd half has no j = 2 tap.
// Rows are linear over the padded width, m = t Fp + f', so that every tap is a constant unit offset for ldmatrix.
#include <cuda_fp16.h>
#include <stdint.h>

#include "mma_util.cuh"
#include "se_internal.h"

namespace se {
namespace {

using namespace mma_util;
constexpr int T = kFramesPerChunk;

template <int CIN, int COUT>
struct DecCfg {
    static constexpr int NH = CIN / 8;                     // channel octets = planes of the input in shared memory
    static constexpr int KS = CIN >= 16 ? 9 * CIN / 16 : 5;  // deconv k-steps (CIN = 8: two taps per step, the 10th is zero)
    static constexpr int NTC = COUT / 8;                   // n-tiles per parity / per kind
    static constexpr int NT = 2 * NTC;
    static constexpr int SLACK = 64;
};

// A fragment of 16 consecutive skip rows straight from global memory (the k order is free as long as A and B agree:
// lane tg supplies channels 4 tg .. 4 tg + 3 of its rows for COUT = 16, channels 2 tg, 2 tg + 1 for COUT = 8)
template <int COUT>
__device__ __forceinline__ void skip_frag(const DecMmaParams& p, const __half* sbase, int tile, int lane, uint32_t (&a)[4],
                                          int (&t)[2], int (&f)[2], bool (&v)[2]) {
    const int g = lane >> 2, tg = lane & 3;
    const int total = T * p.Fs;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = tile * 16 + g + 8 * h;
        v[h] = r < total;
        const int rc = v[h] ? r : total - 1;  // clamped: computed on valid memory, never used
        t[h] = div_magic(rc, p.magic_Fs);
        f[h] = rc - t[h] * p.Fs;
        const __half* src = sbase + (long long)t[h] * p.sk_sT + (long long)f[h] * p.sk_sF;
        if (COUT == 16) {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(src + 4 * tg));
            a[h] = u.x;
            a[2 + h] = u.y;
        } else {
            a[h] = __ldg(reinterpret_cast<const unsigned int*>(src + 2 * tg));
            a[2 + h] = 0u;
        }
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1) dec_mma_kernel(DecMmaParams p) {
    using S = DecCfg<CIN, COUT>;
    constexpr int NH = S::NH, KS = S::KS, NTC = S::NTC, NT = S::NT;
    extern __shared__ __align__(128) unsigned char smem[];
    const int Fp = p.Fp;
    const int plane = p.Tp * Fp + S::SLACK;  // units per octet plane
    unsigned char* sx = smem;
    __half* sy = reinterpret_cast<__half*>(smem + p.off_y);
    uint2* swf = reinterpret_cast<uint2*>(smem + p.off_wf);
    float* spar = reinterpret_cast<float*>(swf + KS * NT * 32);  // deconv bias [2 COUT]
    double* s_red = reinterpret_cast<double*>(spar + 2 * COUT);   // [2][16]
    float* s_co = reinterpret_cast<float*>(s_red + 2 * kWarps);   // mean, inv, mean_r, inv_r
    const uint32_t x_smem = smem_u32(sx);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;

    // ---- one-time set-up -----------------------------------------------------------------------------------------
    for (int i = tid; i < NH * plane; i += kThreads) reinterpret_cast<uint4*>(sx)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < KS * NT * 32; i += kThreads) {  // deconv B fragments: column n = parity * COUT + co
        const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
        const int n = nt * 8 + (ln >> 2), k0 = ks * 16 + 2 * (ln & 3);
        const float* wr = p.w + (long long)n * p.Kp;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + (e & 1) + 8 * (e >> 1);
            v[e] = k < 9 * CIN ? __ldg(wr + k) : 0.f;
        }
        swf[i] = make_uint2(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]));
    }
    for (int i = tid; i < 2 * COUT; i += kThreads) spar[i] = __ldg(p.bias + i);
    // skip pair: B fragments and biases in registers (one k-step).  n-tiles [0, NTC): residualmask, [NTC, 2 NTC): residual
    uint32_t wsk[NT][2];
    float bsk[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int kind = nt / NTC, ch = (nt % NTC) * 8;
        const float* wr = p.w2 + (long long)(2 * (ch + g) + kind) * p.K2p;
        if (COUT == 16) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr) + tg);
            wsk[nt][0] = pack_h2(w4.x, w4.y);
            wsk[nt][1] = pack_h2(w4.z, w4.w);
        } else {
            wsk[nt][0] = pack_h2(__ldg(wr + 2 * tg), __ldg(wr + 2 * tg + 1));
            wsk[nt][1] = 0u;
        }
        bsk[nt][0] = __ldg(p.bias2 + 2 * (ch + 2 * tg) + kind);
        bsk[nt][1] = __ldg(p.bias2 + 2 * (ch + 2 * tg + 1) + kind);
    }

    auto issue_load = [&](int b) {  // global [Tp][Fp][NH] units -> octet planes
        const uint4* src = reinterpret_cast<const uint4*>(p.in + (long long)b * p.in_sB);
        const int total = p.Tp * Fp * NH;
        for (int u = tid; u < total; u += kThreads) {
            const int h = NH == 1 ? 0 : (u & (NH - 1));
            const int pl = NH == 1 ? u : u / NH;
            cp_async16(x_smem + 16u * (h * plane + pl), src + u);
        }
        cp_async_commit();
    };

    // per-lane ldmatrix role: matrices 0/1 = rows 0-7 / 8-15 of the step's first octet (CIN = 8: first tap), 2/3 = second
    const int mi = lane >> 3;
    const int rowoff = (lane & 7) + 8 * (mi & 1);
    const int sel = mi >> 1;
    const int dFp = p.d * Fp;
    const int MT = (T * Fp + 15) >> 4;     // deconv tiles (rows over the padded width)
    int mt_lo, mt_hi;
    warp_tile_range(warp, kWarps, MT, mt_lo, mt_hi);
    const int Fy = 2 * p.Fin - 1, Fs = p.Fs;
    const int MT2 = (T * Fs + 15) >> 4;    // skip tiles
    int st_lo, st_hi;
    warp_tile_range(warp, kWarps, MT2, st_lo, st_hi);
    constexpr int UPR = COUT / 8;

    __syncthreads();
    if (blockIdx.x < p.B) issue_load(p.b0 + blockIdx.x);
    for (int stream = blockIdx.x; stream < p.B; stream += gridDim.x) {
        const int b = p.b0 + stream;
        const __half* sbase = p.skip + (long long)b * p.sk_sB;
        cp_async_wait_all();
        __syncthreads();
        float psum = 0.f, psq = 0.f, rsum = 0.f, rsq = 0.f;
        // ---- pass 1a: transposed conv + ELU -> Y (fp16, [T][Fs][COUT]) + statistics of y --------------------------
        for (int mt = mt_lo; mt < mt_hi; mt += 2) {  // two 16-row tiles per iteration: one B fragment feeds two mma
            const bool two = mt + 1 < mt_hi;
            float acc[2][NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 cb = *reinterpret_cast<const float2*>(spar + 8 * nt + 2 * tg);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    acc[q][nt][0] = acc[q][nt][2] = cb.x;
                    acc[q][nt][1] = acc[q][nt][3] = cb.y;
                }
            }
            const uint32_t abase = x_smem + (uint32_t)((mt * 16 + rowoff) * 16);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                int unit;
                bool odd_live = true;  // the odd output bins have no j = 2 tap (weights are zero there: skip the mma)
                if (CIN == 8) {
                    const int tapA = 2 * ks, tapB = 2 * ks + 1 < 9 ? 2 * ks + 1 : 8;
                    const int uA = (2 - tapA / 3) * dFp + (2 - tapA % 3), uB = (2 - tapB / 3) * dFp + (2 - tapB % 3);
                    unit = sel ? uB : uA;
                } else {
                    const int tap = CIN == 32 ? ks >> 1 : ks;
                    const int oct = CIN == 32 ? 2 * (ks & 1) + sel : sel;
                    unit = (2 - tap / 3) * dFp + (2 - tap % 3) + oct * plane;
                    odd_live = tap % 3 != 2;
                }
                uint32_t a[2][4];
                ldsm_x4(abase + (uint32_t)(unit * 16), a[0]);
                ldsm_x4(abase + (uint32_t)((unit + 16) * 16), a[1]);  // the odd tail reads into the slack, never stored
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (nt < NTC || odd_live) {
                        const uint2 w = swf[(ks * NT + nt) * 32 + lane];
                        mma16816(acc[0][nt], a[0], w.x, w.y);
                        mma16816(acc[1][nt], a[1], w.x, w.y);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 1 && !two) break;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = (mt + q) * 16 + g + 8 * h;
                    const int t = div_magic(r, p.magic_Fp), fi = r - t * Fp;  // input row f' (padded width: fi < Fin is real)
                    const bool rv = t < T && fi < p.Fin;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const int par = nt / NTC, co = (nt % NTC) * 8 + 2 * tg;
                        const int phi = 2 * fi + par;
                        const bool v = rv && phi < Fy;
                        const float e0 = fast_elu(acc[q][nt][2 * h]), e1 = fast_elu(acc[q][nt][2 * h + 1]);
                        if (v) {
                            psum += e0 + e1;
                            psq = fmaf(e0, e0, fmaf(e1, e1, psq));
                            *reinterpret_cast<uint32_t*>(sy + (size_t)(t * Fs + phi) * COUT + co) = pack_h2(e0, e1);
                        }
                    }
                }
            }
        }
        // ---- pass 1b: statistics of the residual-mask pre-activation W_m s + b_m -------------------------------------
        for (int st = st_lo; st < st_hi; st += 2) {
            uint32_t a[2][4];
            int tt[2][2], ff[2][2];
            bool vv[2][2];
            skip_frag<COUT>(p, sbase, st, lane, a[0], tt[0], ff[0], vv[0]);
            const bool two = st + 1 < st_hi;
            if (two) skip_frag<COUT>(p, sbase, st + 1, lane, a[1], tt[1], ff[1], vv[1]);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 1 && !two) break;
#pragma unroll
                for (int nt = 0; nt < NTC; ++nt) {
                    float c[4] = {bsk[nt][0], bsk[nt][1], bsk[nt][0], bsk[nt][1]};
                    mma16816(c, a[q], wsk[nt][0], wsk[nt][1]);
                    if (vv[q][0]) {
                        rsum += c[0] + c[1];
                        rsq = fmaf(c[0], c[0], fmaf(c[1], c[1], rsq));
                    }
                    if (vv[q][1]) {
                        rsum += c[2] + c[3];
                        rsq = fmaf(c[2], c[2], fmaf(c[3], c[3], rsq));
                    }
                }
            }
        }
        __syncthreads();  // X is dead: fetch the next stream's input behind the rest of this one
        if (stream + (int)gridDim.x < p.B) issue_load(b + gridDim.x);
        block_gln(psum, psq, (double)COUT * Fy * T, p.student, s_red, s_co);
        block_gln(rsum, rsq, (double)COUT * Fs * T, p.student, s_red, s_co + 2);
        // ---- pass 2: recompute the skip pair, blend with GLN(y), in place in Y -------------------------------------------
        {
            const float mean = s_co[0], inv = s_co[1], mean_r = s_co[2], inv_r = s_co[3];
            // y * na + nd = GLN(y);  mask = 1 / (1 + 2^(rm * ra + rd))  (sign and log2 e folded into the affine terms)
            float na[NTC][2], nd[NTC][2], ra[NTC][2], rd[NTC][2];
#pragma unroll
            for (int j = 0; j < NTC; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ch = 8 * j + 2 * tg + e;
                    na[j][e] = __ldg(p.nw + ch) * inv;
                    nd[j][e] = fmaf(-mean, na[j][e], __ldg(p.nb + ch));
                    ra[j][e] = __ldg(p.nwr + ch) * inv_r * -kLog2e;
                    rd[j][e] = fmaf(-mean_r, ra[j][e], -kLog2e * __ldg(p.nbr + ch));
                }
            for (int st = st_lo; st < st_hi; st += 2) {
                uint32_t a[2][4];
                int tt[2][2], ff[2][2];
                bool vv[2][2];
                skip_frag<COUT>(p, sbase, st, lane, a[0], tt[0], ff[0], vv[0]);
                const bool two = st + 1 < st_hi;
                if (two) skip_frag<COUT>(p, sbase, st + 1, lane, a[1], tt[1], ff[1], vv[1]);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (q == 1 && !two) break;
#pragma unroll
                    for (int j = 0; j < NTC; ++j) {
                        float cm[4] = {bsk[j][0], bsk[j][1], bsk[j][0], bsk[j][1]};
                        float cr[4] = {bsk[NTC + j][0], bsk[NTC + j][1], bsk[NTC + j][0], bsk[NTC + j][1]};
                        mma16816(cm, a[q], wsk[j][0], wsk[j][1]);
                        mma16816(cr, a[q], wsk[NTC + j][0], wsk[NTC + j][1]);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (!vv[q][h]) continue;
                            __half* yp = sy + (size_t)(tt[q][h] * Fs + ff[q][h]) * COUT + 8 * j + 2 * tg;
                            float y0 = 0.f, y1 = 0.f;
                            if (ff[q][h] < Fy) {
                                const float2 yv = __half22float2(*reinterpret_cast<const __half2*>(yp));
                                y0 = fmaf(yv.x, na[j][0], nd[j][0]);
                                y1 = fmaf(yv.y, na[j][1], nd[j][1]);
                            }
                            // the pre-activation and the residual are rounded to fp16 like the tensors they replace
                            const float2 rm = __half22float2(__floats2half2_rn(cm[2 * h], cm[2 * h + 1]));
                            const float2 rr = __half22float2(__floats2half2_rn(fast_elu(cr[2 * h]), fast_elu(cr[2 * h + 1])));
                            const float m0 = rcp_ftz(1.0f + ex2_ftz(fmaf(rm.x, ra[j][0], rd[j][0])));
                            const float m1 = rcp_ftz(1.0f + ex2_ftz(fmaf(rm.y, ra[j][1], rd[j][1])));
                            *reinterpret_cast<uint32_t*>(yp) = pack_h2(fmaf(m0, rr.x - y0, y0), fmaf(m1, rr.y - y1, y1));
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- pass 3: the block output -> the next block's input interior ---------------------------------------------
        {
            __half* ob = p.out + (long long)b * p.oB;
            const int total = T * Fs * UPR;
            for (int u = tid; u < total; u += kThreads) {
                const int row = u / UPR, c8 = u - row * UPR;
                const int t = div_magic(row, p.magic_Fs), f = row - t * Fs;
                *reinterpret_cast<uint4*>(ob + (long long)t * p.oT + (long long)f * p.oF + 8 * c8) =
                    reinterpret_cast<const uint4*>(sy)[u];
            }
        }
        // Y is rewritten by pass 1a of the next stream only after the barrier at the top of the loop
    }
    cp_async_wait_all();
}

template <int CIN, int COUT>
int launch_dec(DecMmaParams p, cudaStream_t st) {
    using S = DecCfg<CIN, COUT>;
    const int plane = p.Tp * p.Fp + S::SLACK;
    size_t off = (size_t)S::NH * plane * 16;
    p.off_y = (int)off;
    off += (size_t)T * p.Fs * COUT * 2;
    off = (off + 15) / 16 * 16;
    p.off_wf = (int)off;
    off += (size_t)S::KS * S::NT * 32 * 8 + 2 * COUT * 4 + 2 * kWarps * 8 + 32;
    SE_REQUIRE(off <= 227 * 1024, "dec_mma: the stream does not fit in shared memory");
    auto magic = [](int d) { return (uint32_t)(((1ull << 32) + d - 1) / d); };  // exact for dividends < 65536
    SE_REQUIRE(T * p.Fp + 64 < 65536 && T * p.Fs + 64 < 65536, "dec_mma: index range of the magic division");
    p.magic_Fp = magic(p.Fp);
    p.magic_Fs = magic(p.Fs);
    p.magic_in = 0;
    SE_DYN_SMEM((dec_mma_kernel<CIN, COUT>), off);
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return 1;
    dec_mma_kernel<CIN, COUT><<<p.B < num_sms ? p.B : num_sms, kThreads, off, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

bool dec_mma_supported(int Cin, int Cout, int Tp, int Fp, int Fin, int Fs) {
    if (!((Cin == 8 || Cin == 16 || Cin == 32) && (Cout == 8 || Cout == 16))) return false;
    if (2 * Fin - 1 > Fs || Fp != Fin + 2 || Fs < 16) return false;
    const size_t bytes = (size_t)(Cin / 8) * (Tp * Fp + 64) * 16 + (size_t)T * Fs * Cout * 2 + 24 * 1024;
    return bytes <= 227 * 1024;
}

int launch_dec_mma(const DecMmaParams& p, int Cin, int Cout, cudaStream_t st) {
    if (p.B <= 0) return 0;
    SE_REQUIRE(dec_mma_supported(Cin, Cout, p.Tp, p.Fp, p.Fin, p.Fs), "dec_mma: unsupported shape");
    if (Cin == 32 && Cout == 16) return launch_dec<32, 16>(p, st);
    if (Cin == 32 && Cout == 8) return launch_dec<32, 8>(p, st);
    if (Cin == 16 && Cout == 16) return launch_dec<16, 16>(p, st);
    if (Cin == 16 && Cout == 8) return launch_dec<16, 8>(p, st);
    if (Cin == 8 && Cout == 16) return launch_dec<8, 16>(p, st);
    return launch_dec<8, 8>(p, st);
}

}  // namespace se
