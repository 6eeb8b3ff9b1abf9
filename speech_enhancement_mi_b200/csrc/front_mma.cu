// Front end of the CRN chunk step for the small-channel layers (fp16 operand mode), one stream per CTA at a time with
// the stream's activations resident in shared memory:
//
//   preconv3_mma_kernel   the three pre-convolution blocks of CRN_ELU.py:337-339,375-376 (5x5 frequency-dilated causal
//                         conv 5 -> 5 + ELU + gated 1x1 pair + GlobalLayerNorm + residual) in ONE launch.  The feature
//                         tensor is read from HBM once, the three layers update it in place in shared memory, only the
//                         4 carried frames per layer (CRN_ELU.py:243-246) and the final output touch HBM.
//   enc_mma_kernel        one gated causal conv block of the encoder (CRN_ELU.py:230-247, k(5,3), stride (2,1), time
//                         dilation 2^i) for the levels with <= 16 input and <= 32 output channels: conv + ELU + gated 1x1
//                         pair + GlobalLayerNorm in one launch (was: back-to-back tcgen05 GEMM + separate norm pass).
//
// Why not tcgen05 here: with 8 / 16 / 32 output channels a 128-row UMMA tile is A-bandwidth bound (the tensor core
// re-reads 4 KB of A for 16 columns of output) and every frequency / time tap needs its own shifted descriptor; the
// per-tile cost of the issue / commit / TMEM round trip dominated (0.14 - 0.19 ms per layer, 9-16 % of the HBM floor).
// Warp-level mma.sync.m16n8k16 reads each A fragment with one ldmatrix.x4 straight from the resident input (any tap is
// just an address offset), keeps the accumulators in registers where the ELU / gate / statistics epilogue runs, feeds
// the gate GEMM from the accumulator registers (the C fragment of two n-tiles IS the A fragment of a k-step), and needs
// no barriers inside a layer.  GlobalLayerNorm statistics are reduced inside the CTA (no atomics, no second launch).
#include <cuda_fp16.h>
#include <stdint.h>

#include <type_traits>

#include "mma_util.cuh"
#include "se_internal.h"

namespace se {
namespace {

using namespace mma_util;
constexpr int T = kFramesPerChunk;  // 21
constexpr int NB = 201;

// =====================================================================================================================
// three pre-convolution blocks, fused
// =====================================================================================================================
// X[25 frames: 4 carried + 21 new][224 positions][8 halves]: bin f at position 8 + f, zero elsewhere (the conv's
// frequency padding 2 d <= 8 and the overrun of the last 16-row tile).  Output row (t, f), tap (kt, kf) reads unit
// (t + kt) * 224 + 8 + f + (kf - 2) d: sixteen consecutive rows are sixteen consecutive 16-byte units = one conflict-free
// ldmatrix phase per 8 rows.  N = 8 (5 real output channels); a fragment = two frequency taps of one input frame x 8
// channel slots (k = 16).  One warp owns a column of 16 bins and walks the 25 input frames: the three fragments of input
// frame u serve the five output frames u - kt, whose accumulators roll through five register slots.
constexpr int P3_POS = 224;
constexpr int P3_BORDER = 8;
constexpr int P3_ROW_BYTES = P3_POS * 16;
constexpr int P3_X_BYTES = (T + 4) * P3_ROW_BYTES;       // 89,600
constexpr int P3_YPITCH = 208;                           // 13 tiles of 16 bins per frame
constexpr int P3_Y_BYTES = T * P3_YPITCH * 16;           // 69,888: gated values (pre-norm) as fp16 units
constexpr int P3_THREADS = 512;
constexpr int P3_WARPS = P3_THREADS / 32;
// conv B fragments per layer: 10 x (kt, j) with j = frequency-tap pair (0,1) / (2,3) of one input frame, then 3 for tap 4,
// which pairs the SAME tap of two consecutive input frames: (kt 0, kt 1), (kt 2, kt 3), (kt 4, zero)
constexpr int P3_NFR = 13;
constexpr int P3_WF_LAYER = (P3_NFR + 2) * 32;           // uint2 per layer: 13 conv fragments + 2 gate n-tiles
constexpr int P3_WF_BYTES = 3 * P3_WF_LAYER * 8;         // 11,520
constexpr int P3_NPAR = 32;                              // floats per layer: cb[8] bt[8] bg[8] nw[8]
constexpr int P3_PAR_BYTES = 3 * P3_NPAR * 4;
constexpr int P3_OFF_Y = P3_X_BYTES;
constexpr int P3_OFF_WF = P3_OFF_Y + P3_Y_BYTES;
constexpr int P3_OFF_PAR = P3_OFF_WF + P3_WF_BYTES;
constexpr int P3_OFF_RED = P3_OFF_PAR + P3_PAR_BYTES;    // double [2][16]
constexpr int P3_OFF_CO = P3_OFF_RED + 2 * 16 * 8;       // float [2]
constexpr int P3_SMEM = P3_OFF_CO + 16;
static_assert(P3_SMEM <= 227 * 1024, "shared memory budget");

// ldmatrix as a PURE function of (address, epoch): not volatile, so that the compiler may hoist the loads of the next
// input frame above the epilogue of the current one.  `epoch` is produced by a volatile statement behind the barrier
// that published the data, which pins the load below that barrier (and makes loads of different layers distinct).
__device__ __forceinline__ void ldsm_x4_dep(uint32_t addr, uint32_t epoch, uint32_t (&r)[4]) {
    asm("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
        : "r"(addr), "r"(epoch));
}

__global__ void __launch_bounds__(P3_THREADS, 1) preconv3_mma_kernel(Preconv3Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sx = smem;
    unsigned char* sy = smem + P3_OFF_Y;
    uint2* swf = reinterpret_cast<uint2*>(smem + P3_OFF_WF);
    float* spar = reinterpret_cast<float*>(smem + P3_OFF_PAR);
    double* s_red = reinterpret_cast<double*>(smem + P3_OFF_RED);
    float* s_co = reinterpret_cast<float*>(smem + P3_OFF_CO);
    const uint32_t x_smem = smem_u32(sx);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;

    // ---- one-time set-up: B fragments (fp16) and the small fp32 parameters of the three layers ---------------------
    // packed parameter block of a layer (se_internal.h PRECONV_W_*): conv [(kt*5+ci)*28 + kf*5 + co], ...
    // The conv_gated weights and bias carry the factor -log2(e) of the sigmoid (1 / (1 + 2^z)).
    for (int i = tid; i < 3 * P3_WF_LAYER; i += P3_THREADS) {
        const int l = i / P3_WF_LAYER, r = i - l * P3_WF_LAYER, s = r >> 5, ln = r & 31;
        const int gg = ln >> 2, t4 = ln & 3;
        const float* w = l == 0 ? p.w[0] : (l == 1 ? p.w[1] : p.w[2]);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (s < P3_NFR) {  // k 0..7 <- tap (ktA, kfA), k 8..15 <- tap (ktB, kfB); column = output channel gg
            int ktA, kfA, ktB, kfB;
            if (s < 10) {
                ktA = ktB = s >> 1;
                kfA = 2 * (s & 1);
                kfB = kfA + 1;
            } else {
                ktA = 2 * (s - 10);
                ktB = ktA + 1;  // 5: no such tap, zero
                kfA = kfB = 4;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ci = 2 * t4 + e;
                if (ci < 5 && gg < 5) {
                    v[e] = __ldg(w + (ktA * 5 + ci) * 28 + kfA * 5 + gg);
                    if (ktB < 5) v[2 + e] = __ldg(w + (ktB * 5 + ci) * 28 + kfB * 5 + gg);
                }
            }
        } else {  // gate n-tile (0: conv_trans, 1: conv_gated): k = ELU channel 2 t4 + e, column = output channel gg
            const bool gated = s != P3_NFR;
            const int base = gated ? PRECONV_W_WG : PRECONV_W_WT;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 2 * t4 + e;
                if (k < 5 && gg < 5) v[e] = __ldg(w + base + gg * 5 + k) * (gated ? -kLog2e : 1.f);
            }
        }
        swf[i] = make_uint2(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]));
    }
    for (int i = tid; i < 3 * P3_NPAR; i += P3_THREADS) {
        const int l = i / P3_NPAR, r = i - l * P3_NPAR, c = r & 7;
        const float* w = l == 0 ? p.w[0] : (l == 1 ? p.w[1] : p.w[2]);
        float v = 0.f;
        if (c < 5) {
            if (r < 8) v = __ldg(w + PRECONV_W_BIAS + c);
            else if (r < 16) v = __ldg(w + PRECONV_W_BT + c);
            else if (r < 24) v = -kLog2e * __ldg(w + PRECONV_W_BG + c);
            else v = __ldg(w + PRECONV_W_NW + c);
        }
        spar[i] = v;
    }
    __syncthreads();

    // per-lane ldmatrix role: matrices 0/1 = rows 0-7 / 8-15 of the fragment's first tap, 2/3 = of its second tap
    const int mi = lane >> 3;
    const int rowoff = (lane & 7) + 8 * (mi & 1);
    const int tapsel = mi >> 1;
    uint32_t epoch_base = 0;

    for (int stream = blockIdx.x; stream < p.B; stream += gridDim.x) {
        const int b = p.b0 + stream;
        __half* gstate = p.state + (long long)b * p.state_sB;
        {  // frames 0..3 <- carried state of layer 0, frames 4..24 <- this chunk's features
            const uint4* s0 = reinterpret_cast<const uint4*>(gstate);
            const uint4* s1 = reinterpret_cast<const uint4*>(p.feat + (long long)b * p.feat_sB);
            for (int i = tid; i < 4 * P3_POS; i += P3_THREADS) cp_async16(x_smem + 16u * i, s0 + i);
            for (int i = tid; i < T * P3_POS; i += P3_THREADS) cp_async16(x_smem + 4 * P3_ROW_BYTES + 16u * i, s1 + i);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
        }
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const int d = 1 << l;
            uint32_t epoch;  // see ldsm_x4_dep
            asm volatile("mov.u32 %0, %1;" : "=r"(epoch) : "r"(++epoch_base));
            // causal state of this layer for the next chunk: the last 4 frames of its input (CRN_ELU.py:246)
            {
                uint4* dst = reinterpret_cast<uint4*>(gstate + (long long)l * 4 * P3_POS * 8);
                const uint4* src = reinterpret_cast<const uint4*>(sx + T * P3_ROW_BYTES);
                for (int i = tid; i < 4 * P3_POS; i += P3_THREADS) dst[i] = src[i];
            }
            const uint2* wf = swf + l * P3_WF_LAYER;
            const float* par = spar + l * P3_NPAR;
            const float cb0 = par[2 * tg], cb1 = par[2 * tg + 1];
            const float bt0 = par[8 + 2 * tg], bt1 = par[8 + 2 * tg + 1];
            const float bg0 = par[16 + 2 * tg], bg1 = par[16 + 2 * tg + 1];
            const uint2 wgt = wf[P3_NFR * 32 + lane], wgg = wf[(P3_NFR + 1) * 32 + lane];
            uint2 wr[P3_NFR];  // the layer's conv weights stay in registers
#pragma unroll
            for (int i = 0; i < P3_NFR; ++i) wr[i] = wf[i * 32 + lane];
            // byte offset of this lane's tap inside an input frame row, for the three fragments of a frame:
            // j = 0, 1: taps kf = 2 j + tapsel; j = 2: tap 4 of this frame (tapsel 0) / of the next frame (tapsel 1)
            const int o0 = (tapsel - 2) * d * 16, o1 = tapsel * d * 16, o2 = 2 * d * 16 + tapsel * P3_ROW_BYTES;
            float psum = 0.f, psq = 0.f;
            // ---- pass 1.  A warp (0..11) owns a column of 16 bins (the 13th column: see the else branch) and
            // streams over the 25 input frames: the fragments of input frame u feed the output frames u - kt (kt = 0..4),
            // whose accumulators roll through five register slots; output frame u - 4 is complete after frame u: ELU, gate,
            // statistics, gated values -> Y.  3 ldmatrix and 15 mma per 16-row tile (13 + 15 with one fragment per tap
            // pair and no reuse).  The first and last five frames are peeled so that the steady state has no conditions:
            // the compiler interleaves the epilogue of frame u with the mma of frame u + 1 in one basic block.
            if (warp < 12) {
                const int ft = warp;
                // rows 16 ft + g are always real bins (<= 199); rows + 8 run past bin 200 only in the last column
                const float m1 = (ft < 12 || g == 0) ? 1.f : 0.f;
                const uint32_t a_col = x_smem + (uint32_t)((P3_BORDER + 16 * ft + rowoff) * 16);
                unsigned char* y_col = sy + (size_t)((16 * ft + g) * 16 + 4 * tg);
                float acc[5][4];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    acc[i][0] = acc[i][2] = cb0;
                    acc[i][1] = acc[i][3] = cb1;
                }
                // KIND 0: frames 0..4 (outputs u - kt < 0 do not exist), 1: steady state, 2: frames 20..24 (outputs > 20
                // do not exist; the last frame has no successor for the second half of f2)
                auto frames5 = [&](auto kind, int ub) {
                    constexpr int KIND = decltype(kind)::value;
#pragma unroll
                    for (int ui = 0; ui < 5; ++ui) {
                        const int u = ub + ui;
                        uint32_t f0[4], f1[4], f2[4];
                        const uint32_t arow = a_col + (uint32_t)(u * P3_ROW_BYTES);
                        ldsm_x4_dep(arow + o0, epoch, f0);
                        ldsm_x4_dep(arow + o1, epoch, f1);
                        ldsm_x4_dep(arow + ((KIND == 2 && ui == 4) ? 2 * d * 16 : o2), epoch, f2);
#pragma unroll
                        for (int kt = 0; kt < 5; ++kt) {
                            const bool live = KIND == 1 || (KIND == 0 && ui - kt >= 0) || (KIND == 2 && ui - kt <= 0);
                            if (live) {
                                float(&c)[4] = acc[(ui - kt + 5) % 5];
                                mma16816(c, f0, wr[2 * kt].x, wr[2 * kt].y);
                                mma16816(c, f1, wr[2 * kt + 1].x, wr[2 * kt + 1].y);
                                if ((kt & 1) == 0) mma16816(c, f2, wr[10 + kt / 2].x, wr[10 + kt / 2].y);
                            }
                        }
                        if (KIND != 0 || ui == 4) {
                            const int t = u - 4;
                            float(&c)[4] = acc[(ui + 1) % 5];
                            // ELU, then the gated 1x1 pair on the tensor core: the C fragment is the A fragment
                            uint32_t a2[4];
                            a2[0] = pack_h2(fast_elu(c[0]), fast_elu(c[1]));
                            a2[1] = pack_h2(fast_elu(c[2]), fast_elu(c[3]));
                            a2[2] = a2[3] = 0u;
                            float gt[4] = {bt0, bt1, bt0, bt1}, gg4[4] = {bg0, bg1, bg0, bg1};
                            mma16816(gt, a2, wgt.x, wgt.y);
                            mma16816(gg4, a2, wgg.x, wgg.y);
                            // columns >= 5 are exactly 0 (zero weights and biases): every lane stores, no channel mask
                            const float y0 = gt[0] * sigmoid_from_neg_log2(gg4[0]);
                            const float y1 = gt[1] * sigmoid_from_neg_log2(gg4[1]);
                            const float y2 = m1 * gt[2] * sigmoid_from_neg_log2(gg4[2]);
                            const float y3 = m1 * gt[3] * sigmoid_from_neg_log2(gg4[3]);
                            psum += (y0 + y1) + (y2 + y3);
                            psq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, psq))));
                            unsigned char* yr = y_col + (size_t)t * (P3_YPITCH * 16);
                            *reinterpret_cast<uint32_t*>(yr) = pack_h2(y0, y1);
                            *reinterpret_cast<uint32_t*>(yr + 8 * 16) = pack_h2(y2, y3);
                            c[0] = c[2] = cb0;
                            c[1] = c[3] = cb1;
                        }
                    }
                };
                frames5(std::integral_constant<int, 0>{}, 0);
#pragma unroll 1
                for (int ub = 5; ub < 20; ub += 5) frames5(std::integral_constant<int, 1>{}, ub);
                frames5(std::integral_constant<int, 2>{}, 20);
            } else {
                // The 13th column (bins 192..200, 9 real rows) is shared by warps 12..15, a quarter of the output frames
                // each.  The tensor pipe of an SM sub-partition is the bound of this pass (one HMMA per 16 cycles) and
                // warp w issues on sub-partition w % 4: twelve full columns are three per sub-partition, and the four
                // quarters (5 or 6 output frames + 4 frames of run-in each) add 0.4 of a column to every one of them
                // instead of a whole fourth column to sub-partition 0.
                const int q = warp - 12;
                const int ta = 5 * q, tb = q == 3 ? T : 5 * q + 5;  // output frames [ta, tb)
                constexpr int ft = 12;
                const float m1 = g == 0 ? 1.f : 0.f;  // rows + 8 run past bin 200 except for g = 0
                const uint32_t a_col = x_smem + (uint32_t)((P3_BORDER + 16 * ft + rowoff) * 16);
                unsigned char* y_col = sy + (size_t)((16 * ft + g) * 16 + 4 * tg);
                float acc[5][4];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    acc[i][0] = acc[i][2] = cb0;
                    acc[i][1] = acc[i][3] = cb1;
                }
                // input frames ta .. tb + 3 in blocks of five (slot of output frame u - kt: (ui - kt) mod 5 as above);
                // an output frame is live when it lies in [ta, tb)
#pragma unroll 1
                for (int ub = ta; ub < tb + 4; ub += 5) {
                    const bool first = ub == ta;
#pragma unroll
                    for (int ui = 0; ui < 5; ++ui) {
                        const int u = ub + ui;
                        if (u < tb + 4) {  // warp-uniform
                            uint32_t f0[4], f1[4], f2[4];
                            const uint32_t arow = a_col + (uint32_t)(u * P3_ROW_BYTES);
                            ldsm_x4_dep(arow + o0, epoch, f0);
                            ldsm_x4_dep(arow + o1, epoch, f1);
                            ldsm_x4_dep(arow + (u == T + 3 ? 2 * d * 16 : o2), epoch, f2);  // frame 24 has no successor
#pragma unroll
                            for (int kt = 0; kt < 5; ++kt) {
                                const bool live = (!first || ui >= kt) && (u - kt < tb);
                                if (live) {
                                    float(&c)[4] = acc[(ui - kt + 5) % 5];
                                    mma16816(c, f0, wr[2 * kt].x, wr[2 * kt].y);
                                    mma16816(c, f1, wr[2 * kt + 1].x, wr[2 * kt + 1].y);
                                    if ((kt & 1) == 0) mma16816(c, f2, wr[10 + kt / 2].x, wr[10 + kt / 2].y);
                                }
                            }
                            if ((!first || ui == 4) && u - 4 < tb) {
                                const int t = u - 4;
                                float(&c)[4] = acc[(ui + 1) % 5];
                                uint32_t a2[4];
                                a2[0] = pack_h2(fast_elu(c[0]), fast_elu(c[1]));
                                a2[1] = pack_h2(fast_elu(c[2]), fast_elu(c[3]));
                                a2[2] = a2[3] = 0u;
                                float gt[4] = {bt0, bt1, bt0, bt1}, gg4[4] = {bg0, bg1, bg0, bg1};
                                mma16816(gt, a2, wgt.x, wgt.y);
                                mma16816(gg4, a2, wgg.x, wgg.y);
                                const float y0 = gt[0] * sigmoid_from_neg_log2(gg4[0]);
                                const float y1 = gt[1] * sigmoid_from_neg_log2(gg4[1]);
                                const float y2 = m1 * gt[2] * sigmoid_from_neg_log2(gg4[2]);
                                const float y3 = m1 * gt[3] * sigmoid_from_neg_log2(gg4[3]);
                                psum += (y0 + y1) + (y2 + y3);
                                psq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, psq))));
                                unsigned char* yr = y_col + (size_t)t * (P3_YPITCH * 16);
                                *reinterpret_cast<uint32_t*>(yr) = pack_h2(y0, y1);
                                *reinterpret_cast<uint32_t*>(yr + 8 * 16) = pack_h2(y2, y3);
                                c[0] = c[2] = cb0;
                                c[1] = c[3] = cb1;
                            }
                        }
                    }
                }
            }
            __syncthreads();  // frames 0..3 of X are dead from here on
            if (l < 2) {      // carried state of the next layer -> frames 0..3 (overlaps the statistics and pass 2)
                const uint4* s0 = reinterpret_cast<const uint4*>(gstate + (long long)(l + 1) * 4 * P3_POS * 8);
                for (int i = tid; i < 4 * P3_POS; i += P3_THREADS) cp_async16(x_smem + 16u * i, s0 + i);
                cp_async_commit();
            }
            block_gln<P3_WARPS>(psum, psq, 5.0 * NB * T, p.student, s_red, s_co);
            // ---- pass 2: normalise + residual (CRN_ELU.py:376), in place (next layer's input) or to the output ----------
            {
                const float mean = s_co[0], inv = s_co[1];
                float na[5], nd[5];  // y * na + nd = (y - mean) * inv * w + b
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    na[c] = par[24 + c] * inv;
                    nd[c] = fmaf(-mean, na[c], __ldg((l == 0 ? p.w[0] : (l == 1 ? p.w[1] : p.w[2])) + PRECONV_W_NB + c));
                }
                __half* ob = p.out + (long long)b * p.oB;
                for (int i = tid; i < T * NB; i += P3_THREADS) {
                    const int t = i / NB, f = i - t * NB;
                    uint4* xu = reinterpret_cast<uint4*>(sx + ((t + 4) * P3_POS + P3_BORDER + f) * 16);
                    float yv[8], xv[8];
                    unpack8(*reinterpret_cast<const uint4*>(sy + (t * P3_YPITCH + f) * 16), yv);
                    unpack8(*xu, xv);
                    float o[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) o[c] = fmaf(yv[c], na[c], nd[c]) + xv[c];
                    uint4 u;
                    u.x = pack_h2(o[0], o[1]);
                    u.y = pack_h2(o[2], o[3]);
                    u.z = pack_h2(o[4], 0.f);
                    u.w = 0u;
                    if (l < 2) *xu = u;
                    else *reinterpret_cast<uint4*>(ob + (long long)t * p.oT + (long long)f * p.oF) = u;
                }
            }
            cp_async_wait_all();
            __syncthreads();
        }
    }
}

// =====================================================================================================================
// encoder block with few channels: conv k(5,3) stride (2,1) dilation (1,dt) + ELU + gated 1x1 pair + GlobalLayerNorm
// =====================================================================================================================
// The stream's zero-bordered input [Tp][Fp][CIN] is de-interleaved on the way into shared memory into planes
// X[parity of the padded bin][channel octet][Tp][Jp] of 16-byte units (Jp = ceil(Fp / 2)), so that the stride-2 rows of
// a tap are consecutive units: output row m = t Jp + f' (f' < Fo valid, the 2 extra rows per frame are computed and
// dropped) and tap (kt, kf) read unit m + kt dt Jp + (kf >> 1) of plane (kf & 1, octet).
template <int CIN, int COUT>
struct EncCfg {
    static constexpr int NH = CIN / 8;
    static constexpr int KS = (15 * CIN + 15) / 16;  // conv k-steps (CIN = 8: two taps per step, the 16th tap is zero)
    static constexpr int NT = COUT / 8;
    static constexpr int KS2 = (COUT + 15) / 16;
    static constexpr int NT2 = 2 * NT;
    static constexpr int SLACK = 64;  // units behind every plane (tile overrun of the last frame)
};

// NQ consecutive 16-row tiles starting at mt0 (all < MT): conv, ELU, gate, statistics, gated values -> Y
template <int CIN, int COUT, int NQ>
__device__ __forceinline__ void enc_tiles(int mt0, const EncMmaParams& p, uint32_t x_smem, int rowoff, int sel, int dtJ,
                                          int plane, int Jp, const uint2* swf, const uint2* swf2, const float* spar,
                                          unsigned char* sy, int lane, float& psum, float& psq) {
    using S = EncCfg<CIN, COUT>;
    constexpr int NH = S::NH, KS = S::KS, NT = S::NT, KS2 = S::KS2, NT2 = S::NT2;
    const int g = lane >> 2, tg = lane & 3;
    float acc[NQ][NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float2 cb = *reinterpret_cast<const float2*>(spar + 8 * nt + 2 * tg);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            acc[q][nt][0] = acc[q][nt][2] = cb.x;
            acc[q][nt][1] = acc[q][nt][3] = cb.y;
        }
    }
    const uint32_t abase = x_smem + (uint32_t)((mt0 * 16 + rowoff) * 16);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        int unit;
        if (CIN == 8) {
            const int tapA = 2 * ks, tapB = 2 * ks + 1 < 15 ? 2 * ks + 1 : 14;  // tap 15: zero weights, valid address
            const int uA = (tapA / 5) * dtJ + ((tapA % 5) >> 1) + ((tapA % 5) & 1) * plane;
            const int uB = (tapB / 5) * dtJ + ((tapB % 5) >> 1) + ((tapB % 5) & 1) * plane;
            unit = sel ? uB : uA;
        } else {
            const int kt = ks / 5, kf = ks % 5;
            unit = kt * dtJ + (kf >> 1) + ((kf & 1) * NH + sel) * plane;
        }
        uint32_t a[NQ][4];
#pragma unroll
        for (int q = 0; q < NQ; ++q) ldsm_x4(abase + (uint32_t)((unit + 16 * q) * 16), a[q]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint2 w = swf[(ks * NT + nt) * 32 + lane];
#pragma unroll
            for (int q = 0; q < NQ; ++q) mma16816(acc[q][nt], a[q], w.x, w.y);
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        uint32_t a2[KS2][4];
#pragma unroll
        for (int s = 0; s < KS2; ++s) {
            a2[s][0] = pack_h2(fast_elu(acc[q][2 * s][0]), fast_elu(acc[q][2 * s][1]));
            a2[s][1] = pack_h2(fast_elu(acc[q][2 * s][2]), fast_elu(acc[q][2 * s][3]));
            if (2 * s + 1 < NT) {
                a2[s][2] = pack_h2(fast_elu(acc[q][2 * s + 1][0]), fast_elu(acc[q][2 * s + 1][1]));
                a2[s][3] = pack_h2(fast_elu(acc[q][2 * s + 1][2]), fast_elu(acc[q][2 * s + 1][3]));
            } else {
                a2[s][2] = a2[s][3] = 0u;
            }
        }
        const int r0 = (mt0 + q) * 16 + g, r1 = r0 + 8;
        const int t0 = div_magic(r0, p.magic_Jp), f0 = r0 - t0 * Jp, t1 = div_magic(r1, p.magic_Jp), f1 = r1 - t1 * Jp;
        const bool v0 = t0 < T && f0 < p.Fo, v1 = t1 < T && f1 < p.Fo;
        const float m0 = v0 ? 1.f : 0.f, m1 = v1 ? 1.f : 0.f;
        __half* y0 = reinterpret_cast<__half*>(sy) + (size_t)(t0 * p.Fo + f0) * COUT + 2 * tg;
        __half* y1 = reinterpret_cast<__half*>(sy) + (size_t)(t1 * p.Fo + f1) * COUT + 2 * tg;
#pragma unroll
        for (int j = 0; j < NT; ++j) {  // trans tile j and gated tile NT + j land in the same lanes and slots
            const float2 bt = *reinterpret_cast<const float2*>(spar + COUT + 8 * j + 2 * tg);
            const float2 bg = *reinterpret_cast<const float2*>(spar + 2 * COUT + 8 * j + 2 * tg);
            float gt[4] = {bt.x, bt.y, bt.x, bt.y}, gg4[4] = {bg.x, bg.y, bg.x, bg.y};
#pragma unroll
            for (int s = 0; s < KS2; ++s) {
                const uint2 wt = swf2[(s * NT2 + j) * 32 + lane], wg = swf2[(s * NT2 + NT + j) * 32 + lane];
                mma16816(gt, a2[s], wt.x, wt.y);
                mma16816(gg4, a2[s], wg.x, wg.y);
            }
            const float ya = m0 * gt[0] * sigmoid_from_neg_log2(gg4[0]), yb = m0 * gt[1] * sigmoid_from_neg_log2(gg4[1]);
            const float yc = m1 * gt[2] * sigmoid_from_neg_log2(gg4[2]), yd = m1 * gt[3] * sigmoid_from_neg_log2(gg4[3]);
            psum += (ya + yb) + (yc + yd);
            psq = fmaf(ya, ya, fmaf(yb, yb, fmaf(yc, yc, fmaf(yd, yd, psq))));
            if (v0) *reinterpret_cast<uint32_t*>(y0 + 8 * j) = pack_h2(ya, yb);
            if (v1) *reinterpret_cast<uint32_t*>(y1 + 8 * j) = pack_h2(yc, yd);
        }
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1) enc_mma_kernel(EncMmaParams p) {
    using S = EncCfg<CIN, COUT>;
    constexpr int NH = S::NH, KS = S::KS, NT = S::NT, KS2 = S::KS2, NT2 = S::NT2;
    extern __shared__ __align__(128) unsigned char smem[];
    const int Jp = (p.Fp + 1) >> 1;
    const int plane = p.Tp * Jp + S::SLACK;  // units per plane
    unsigned char* sx = smem;
    unsigned char* sy = smem + p.off_y;
    uint2* swf = reinterpret_cast<uint2*>(smem + p.off_wf);
    uint2* swf2 = swf + KS * NT * 32;
    float* spar = reinterpret_cast<float*>(swf2 + KS2 * NT2 * 32);  // cb[COUT] | b2t[COUT] | -log2(e) b2g[COUT]
    double* s_red = reinterpret_cast<double*>(spar + 3 * COUT);
    float* s_co = reinterpret_cast<float*>(s_red + 2 * kWarps);
    const uint32_t x_smem = smem_u32(sx);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time set-up -----------------------------------------------------------------------------------------
    for (int i = tid; i < 2 * NH * plane; i += kThreads) reinterpret_cast<uint4*>(sx)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < KS * NT * 32; i += kThreads) {  // conv B fragments: column n = output channel, k = (tap, ci)
        const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
        const int n = nt * 8 + (ln >> 2), k0 = ks * 16 + 2 * (ln & 3);
        const float* wr = p.w + (long long)n * p.Kp;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + (e & 1) + 8 * (e >> 1);
            v[e] = k < 15 * CIN ? __ldg(wr + k) : 0.f;
        }
        swf[i] = make_uint2(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]));
    }
    // gate B fragments: n-tiles [0,NT) conv_trans, [NT,2NT) conv_gated scaled by -log2(e) (sigmoid = 1 / (1 + 2^z))
    for (int i = tid; i < KS2 * NT2 * 32; i += kThreads) {
        const int ln = i & 31, nt2 = (i >> 5) % NT2, ks2 = (i >> 5) / NT2;
        const int kind = nt2 / NT, ch = (nt2 % NT) * 8 + (ln >> 2), k0 = ks2 * 16 + 2 * (ln & 3);
        const float* wr = p.w2 + (long long)(2 * ch + kind) * p.w2_pitch;
        const float sc = kind ? -kLog2e : 1.f;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + (e & 1) + 8 * (e >> 1);
            v[e] = k < COUT ? sc * __ldg(wr + k) : 0.f;
        }
        swf2[i] = make_uint2(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]));
    }
    for (int i = tid; i < COUT; i += kThreads) {
        spar[i] = __ldg(p.bias + i);
        spar[COUT + i] = __ldg(p.bias2 + 2 * i);
        spar[2 * COUT + i] = -kLog2e * __ldg(p.bias2 + 2 * i + 1);
    }

    auto issue_load = [&](int b) {  // global [Tp][Fp][NH] units -> planes
        const uint4* src = reinterpret_cast<const uint4*>(p.in + (long long)b * p.in_sB);
        const int total = p.Tp * p.Fp * NH;
        for (int u = tid; u < total; u += kThreads) {
            const int h = NH == 1 ? 0 : (u & (NH - 1));
            const int pl = NH == 1 ? u : u / NH;
            const int tt = div_magic(pl, p.magic_Fp), pos = pl - tt * p.Fp;
            const int unit = ((pos & 1) * NH + h) * plane + tt * Jp + (pos >> 1);
            cp_async16(x_smem + 16u * unit, src + u);
        }
        cp_async_commit();
    };

    // per-lane ldmatrix role.  CIN = 8: matrices 0/1 = rows 0-7 / 8-15 of the step's first tap, 2/3 = of its second tap;
    // CIN = 16: matrices 0/1 = channel octet 0, 2/3 = octet 1 of the step's single tap
    const int mi = lane >> 3;
    const int rowoff = (lane & 7) + 8 * (mi & 1);
    const int sel = mi >> 1;
    const int dtJ = p.dt * Jp;
    const int MT = (T * Jp + 15) >> 4;
    int mt_lo, mt_hi;  // the warp's contiguous tile range
    warp_tile_range(warp, kWarps, MT, mt_lo, mt_hi);
    const int Fo = p.Fo;
    const double count = (double)COUT * Fo * T;
    constexpr int UPR = COUT / 8;  // 16-byte units per output row

    __syncthreads();
    if (blockIdx.x < p.B) issue_load(p.b0 + blockIdx.x);
    for (int stream = blockIdx.x; stream < p.B; stream += gridDim.x) {
        const int b = p.b0 + stream;
        cp_async_wait_all();
        __syncthreads();
        float psum = 0.f, psq = 0.f;
        // ---- pass 1: conv + ELU + gate -> Y (fp16, [T][Fo][COUT]) + statistics; two 16-row tiles in flight ------------
        int mt = mt_lo;
        for (; mt + 2 <= mt_hi; mt += 2)
            enc_tiles<CIN, COUT, 2>(mt, p, x_smem, rowoff, sel, dtJ, plane, Jp, swf, swf2, spar, sy, lane, psum, psq);
        if (mt < mt_hi)
            enc_tiles<CIN, COUT, 1>(mt, p, x_smem, rowoff, sel, dtJ, plane, Jp, swf, swf2, spar, sy, lane, psum, psq);
        __syncthreads();  // X is dead: fetch the next stream's input while this one is normalised and written out
        if (stream + (int)gridDim.x < p.B) issue_load(b + gridDim.x);
        block_gln(psum, psq, count, p.student, s_red, s_co);
        // ---- pass 2: GlobalLayerNorm -> the next block's input interior --------------------------------------------
        {
            const float mean = s_co[0], inv = s_co[1];
            const int c8 = tid % UPR;  // kThreads % UPR == 0: a thread always serves the same channel octet
            float na[8], nd[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                na[k] = __ldg(p.nw + 8 * c8 + k) * inv;
                nd[k] = fmaf(-mean, na[k], __ldg(p.nb + 8 * c8 + k));
            }
            __half* ob = p.out + (long long)b * p.oB + 8 * c8;
            const int total = T * Fo * UPR;
            for (int u = tid; u < total; u += kThreads) {
                const int row = u / UPR;
                const int t = div_magic(row, p.magic_Fo), f = row - t * Fo;
                float v[8];
                unpack8(reinterpret_cast<const uint4*>(sy)[u], v);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], na[k], nd[k]);
                *reinterpret_cast<uint4*>(ob + (long long)t * p.oT + (long long)f * p.oF) = pack8(v);
            }
        }
    }
    cp_async_wait_all();
}

template <int CIN, int COUT>
int launch_enc(EncMmaParams p, cudaStream_t st) {
    using S = EncCfg<CIN, COUT>;
    const int Jp = (p.Fp + 1) / 2;
    const int plane = p.Tp * Jp + S::SLACK;
    size_t off = (size_t)2 * S::NH * plane * 16;
    p.off_y = (int)off;
    off += (size_t)T * p.Fo * COUT * 2;
    off = (off + 15) / 16 * 16;
    p.off_wf = (int)off;
    off += (size_t)(S::KS * S::NT + S::KS2 * S::NT2) * 32 * 8 + 3 * COUT * 4 + 2 * kWarps * 8 + 16;
    SE_REQUIRE(off <= 227 * 1024, "enc_mma: the stream does not fit in shared memory");
    auto magic = [](int d) { return (uint32_t)(((1ull << 32) + d - 1) / d); };  // exact for dividends < 65536
    SE_REQUIRE(p.Tp * p.Fp < 65536 && T * Jp + 64 < 65536, "enc_mma: index range of the magic division");
    p.magic_Jp = magic(Jp);
    p.magic_Fo = magic(p.Fo);
    p.magic_Fp = magic(p.Fp);
    SE_DYN_SMEM((enc_mma_kernel<CIN, COUT>), off);
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return 1;
    enc_mma_kernel<CIN, COUT><<<p.B < num_sms ? p.B : num_sms, kThreads, off, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

bool enc_mma_supported(int Cin, int Cout, int Tp, int Fp, int Fo) {
    if (!((Cin == 8 || Cin == 16) && (Cout == 8 || Cout == 16 || Cout == 32))) return false;
    const int Jp = (Fp + 1) / 2;
    const size_t bytes = (size_t)2 * (Cin / 8) * (Tp * Jp + 64) * 16 + (size_t)T * Fo * Cout * 2 + 40 * 1024;
    return bytes <= 227 * 1024 && Fo <= Jp;
}

int launch_enc_mma(const EncMmaParams& p, int Cin, int Cout, cudaStream_t st) {
    if (p.B <= 0) return 0;
    SE_REQUIRE(enc_mma_supported(Cin, Cout, p.Tp, p.Fp, p.Fo), "enc_mma: unsupported shape");
    if (Cin == 8 && Cout == 8) return launch_enc<8, 8>(p, st);
    if (Cin == 8 && Cout == 16) return launch_enc<8, 16>(p, st);
    if (Cin == 8 && Cout == 32) return launch_enc<8, 32>(p, st);
    if (Cin == 16 && Cout == 8) return launch_enc<16, 8>(p, st);
    if (Cin == 16 && Cout == 16) return launch_enc<16, 16>(p, st);
    return launch_enc<16, 32>(p, st);
}

int launch_preconv3(const Preconv3Params& p, cudaStream_t st) {
    if (p.B <= 0) return 0;
    SE_DYN_SMEM(preconv3_mma_kernel, P3_SMEM);
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return 1;
    preconv3_mma_kernel<<<p.B < num_sms ? p.B : num_sms, P3_THREADS, P3_SMEM, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace se
