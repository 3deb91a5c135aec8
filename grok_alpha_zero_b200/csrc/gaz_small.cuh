// gaz_small.cuh -- the layers of the policy/value networks that are too small for a tcgen05 pipeline, on the warp-level
// tensor cores (mma.sync.m16n8k16, bf16 x bf16 -> fp32) or fused on the CUDA cores:
//
//   stem_mma_kernel      first convolution on the 2 / 4 input planes + BN + relu / gelu (*/Build_Model.py "eyes" layers):
//                        a board is a small implicit GEMM, A = im2col fragments read straight from the zero-bordered input
//                        planes in shared memory (entries -1 / 0 / +1, exact in bf16), B = the filters split into bf16
//                        hi + lo parts (fp32 filter accuracy, two MMAs per fragment).  Outputs: the fp32 residual stream
//                        (blocked layout, gaz_conv::f32_blk_index) and / or the bf16 operands of the first block.
//   headconv_wide_kernel head convolutions that read the fp32 trunk output directly (Connect4: 3x3 C128 -> C8 for the policy
//                        AND the value head, Connect4/Build_Model.py:27,48; TicTacToe: 1x1 C64 -> C8 / C4): both heads in
//                        one pass (N = 16), activations split into bf16 hi + lo while they are staged in shared memory and
//                        weights split the same way, three MMAs per fragment (hi*hi + lo*hi + hi*lo): fp32-level accuracy
//                        at tensor-core speed.  The CUDA-core form of this layer was 19 % of a Connect4 search round.
//   mlp_chain_kernel     the dense stack behind a head (Connect4: 336 -> 128 -> 64 -> 7 / 1 with BN + ReLU in between) as
//                        ONE launch per head: a CTA keeps 16 leaves' activations in shared memory between the layers.
//                        Six 20 us latency-bound SGEMM launches were 11 % of a Connect4 round.
#pragma once
#include "gaz_conv.cuh"
#include <cuda_bf16.h>

namespace gaz_small {
using gaz_conv::f32_blk_index;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) (Keras Activation("gelu"), exact form).  erf by Abramowitz-Stegun 7.1.26 (branch-free,
// |error| <= 1.5e-7 absolute - below fp32 round-off of the result for |x| >= 1 and far below the bf16 rounding the outputs get):
// libdevice's erff costs ~3x the instructions and made the stem instruction-bound (29 M evaluations per Connect4 round).
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    float pl = fmaf(1.061405429f, t, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    const float e = 1.0f - pl * t * __expf(-z * z);     // erf(|x| / sqrt 2)
    return 0.5f * x * (1.0f + copysignf(e, x));
}

// B fragment of mma.m16n8k16 (k x n, "col"): lane (g = lane >> 2, t = lane & 3) holds k = 2t, 2t+1 (register 0) and
// k = 2t+8, 2t+9 (register 1) of column n = g.  Host helper: bf16 hi / lo split of w[k][n] pairs -> {hi0, hi1, lo0, lo1}.
struct Frag { uint32_t hi0, hi1, lo0, lo1; };
static inline uint16_t host_bf16(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16); }
static inline float host_bf16_val(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
template <class F> static inline Frag host_frag(int lane, F w /* (k, n) -> float, 0 outside */, int k0, int n0) {
    const int g = lane >> 2, t = lane & 3;
    uint16_t hi[4], lo[4];
    const int ks[4] = {k0 + 2 * t, k0 + 2 * t + 1, k0 + 2 * t + 8, k0 + 2 * t + 9};
    for (int i = 0; i < 4; i++) {
        const float v = w(ks[i], n0 + g);
        hi[i] = host_bf16(v);
        lo[i] = host_bf16(v - host_bf16_val(hi[i]));
    }
    Frag f;
    f.hi0 = (uint32_t)hi[0] | ((uint32_t)hi[1] << 16); f.hi1 = (uint32_t)hi[2] | ((uint32_t)hi[3] << 16);
    f.lo0 = (uint32_t)lo[0] | ((uint32_t)lo[1] << 16); f.lo1 = (uint32_t)lo[2] | ((uint32_t)lo[3] << 16);
    return f;
}

// ------------------------------------------------------------------------------------------------ stem ----
struct StemMmaArgs {
    const int32_t *count;
    int max_count;
    const int8_t *states;   // [leaf][H*W*Cin] HWC
    int H, W, Cout, P_pad, Wp, act;   // act: GAZ_ACT_* of the stem
    int G;                  // boards staged per CTA iteration (G * P_pad is a multiple of 16)
    const uint4 *frags;     // [k-step][n-tile][lane] {hi0, hi1, lo0, lo1}, k = (ky*K + kx)*Cin + ci, zero beyond K*K*Cin
    const float *par;       // [5][Cout]: conv bias | BN scale | BN shift | scale_a | shift_a
    __nv_bfloat16 *out_q;   // activation itself (bf16 rows), optional
    float *out_raw;         // activation itself (fp32, blocked), optional
    __nv_bfloat16 *out_a;   // relu(scale_a * activation + shift_a) (bf16 rows), optional
};

template <int K, int CIN> struct StemMmaCfg {
    static constexpr int KK = K * K * CIN, KS = (KK + 15) / 16, kh = K >> 1;
    static size_t smem(int H, int W, int Cout, int G) {
        return (size_t)KS * (Cout >> 3) * 32 * 16 + (size_t)5 * Cout * 4 + (size_t)G * (H + 2 * kh) * (W + 2 * kh) * CIN * 2 + 16;
    }
};

template <int K, int CIN> __global__ void __launch_bounds__(256, 2) stem_mma_kernel(StemMmaArgs p) {
    using Cfg = StemMmaCfg<K, CIN>;
    constexpr int KS = Cfg::KS, kh = Cfg::kh, KK = Cfg::KK;
    extern __shared__ __align__(16) uint8_t smem[];
    const int NT = p.Cout >> 3;
    uint4 *s_frag = reinterpret_cast<uint4 *>(smem);
    float *s_par = reinterpret_cast<float *>(s_frag + KS * NT * 32);
    __nv_bfloat16 *s_in = reinterpret_cast<__nv_bfloat16 *>(s_par + 5 * p.Cout);
    const int WPc = p.W + 2 * kh, pitch = WPc * CIN, board_elems = (p.H + 2 * kh) * pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) s_frag[i] = p.frags[i];
    for (int i = threadIdx.x; i < 5 * p.Cout; i += blockDim.x) s_par[i] = p.par[i];
    for (int i = threadIdx.x; i < p.G * board_elems; i += blockDim.x) s_in[i] = __float2bfloat16_rn(0.0f);
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const long long total_rows = ((long long)cnt * p.P_pad + 255) / 256 * 256;   // rows the tcgen05 tiles of the next layer read
    const int n_boards_pad = (int)((total_rows + p.P_pad - 1) / p.P_pad);
    const int n_groups = (n_boards_pad + p.G - 1) / p.G;
    const int nin = p.H * p.W * CIN;
    const int n_mt = p.G * p.P_pad / 16;
    // element offset (inside a zero-bordered board) of this lane's k pairs, per k-step: k = ks*16 + 2t (+8)
    int koff[KS][2];
#pragma unroll
    for (int ks = 0; ks < KS; ks++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int k = ks * 16 + 2 * t + 8 * h;
            const int tap = k / CIN, ci = k - tap * CIN, ky = tap / K, kx = tap - ky * K;
            koff[ks][h] = k < KK ? (ky * WPc + kx) * CIN + ci : 0;   // beyond K*K*Cin the filters are zero: any valid address
        }
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < p.G * nin; i += blockDim.x) {
            const int gb = i / nin, e = i - gb * nin;
            const int b = grp * p.G + gb;
            const int cell = e / CIN, ci = e - cell * CIN, y = cell / p.W, x = cell - y * p.W;
            const float v = b < cnt ? (float)p.states[(size_t)b * nin + e] : 0.0f;
            s_in[gb * board_elems + (y + kh) * pitch + (x + kh) * CIN + ci] = __float2bfloat16_rn(v);
        }
        __syncthreads();
        for (int mt = warp; mt < n_mt; mt += (int)(blockDim.x >> 5)) {
            // the two rows of this lane: mt*16 + g and + 8 inside the group
            int base[2];
            bool live[2];
            long long grow[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int rr = mt * 16 + g + 8 * h;
                const int gb = rr / p.P_pad, pos = rr - gb * p.P_pad;
                const int yy = pos / p.Wp - 1, xx = pos % p.Wp;
                live[h] = yy >= 0 && yy < p.H && xx < p.W && (grp * p.G + gb) < cnt;
                base[h] = gb * board_elems + (live[h] ? (yy * WPc + xx) * CIN : 0);
                grow[h] = (long long)(grp * p.G) * p.P_pad + rr;
            }
            const bool any_live[2] = {__any_sync(0xffffffffu, live[0]) != 0, __any_sync(0xffffffffu, live[1]) != 0};
            // blocked fp32 layout (gaz_conv::f32_blk_index): channel 8*nt + 2t of row r sits at rbase + (nt >> 2) * 1024 + (nt & 3) * 256
            const size_t rbase[2] = {f32_blk_index(grow[0], 2 * t, p.Cout), f32_blk_index(grow[1], 2 * t, p.Cout)};
            for (int nc = 0; nc < NT; nc += 8) {     // 8 n-tiles (64 channels) of accumulators at a time: two CTAs per SM
                float acc[8][4];
#pragma unroll
                for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
#pragma unroll
                for (int ks = 0; ks < KS; ks++) {
                    const uint32_t a0 = *reinterpret_cast<const uint32_t *>(s_in + base[0] + koff[ks][0]);
                    const uint32_t a1 = *reinterpret_cast<const uint32_t *>(s_in + base[1] + koff[ks][0]);
                    const uint32_t a2 = *reinterpret_cast<const uint32_t *>(s_in + base[0] + koff[ks][1]);
                    const uint32_t a3 = *reinterpret_cast<const uint32_t *>(s_in + base[1] + koff[ks][1]);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (nc + j < NT) {
                            const uint4 bf = s_frag[(ks * NT + nc + j) * 32 + lane];
                            mma_bf16_16816(acc[j], a0, a1, a2, a3, bf.x, bf.y);
                            mma_bf16_16816(acc[j], a0, a1, a2, a3, bf.z, bf.w);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (nc + j >= NT) continue;
                    const int c = (nc + j) * 8 + 2 * t;
                    const float2 bias = *reinterpret_cast<const float2 *>(s_par + c), sc = *reinterpret_cast<const float2 *>(s_par + p.Cout + c),
                                 sh = *reinterpret_cast<const float2 *>(s_par + 2 * p.Cout + c);
                    const float2 sa = *reinterpret_cast<const float2 *>(s_par + 3 * p.Cout + c), ta = *reinterpret_cast<const float2 *>(s_par + 4 * p.Cout + c);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (grow[h] >= total_rows) continue;
                        if (!any_live[h]) {     // a whole 8-row group of padding (warp-uniform): zeros, no arithmetic
                            if (p.out_raw) *reinterpret_cast<float2 *>(p.out_raw + rbase[h] + (size_t)((((nc + j) >> 2) << 10) + (((nc + j) & 3) << 8))) = make_float2(0.0f, 0.0f);
                            if (p.out_q) *reinterpret_cast<uint32_t *>(p.out_q + (size_t)grow[h] * p.Cout + c) = 0u;
                            if (p.out_a) *reinterpret_cast<uint32_t *>(p.out_a + (size_t)grow[h] * p.Cout + c) = 0u;
                            continue;
                        }
                        float v0 = fmaf(sc.x, acc[j][2 * h] + bias.x, sh.x), v1 = fmaf(sc.y, acc[j][2 * h + 1] + bias.y, sh.y);
                        if (p.act == 1) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                        else if (p.act == 2) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
                        if (!live[h]) { v0 = 0.0f; v1 = 0.0f; }
                        if (p.out_raw) *reinterpret_cast<float2 *>(p.out_raw + rbase[h] + (size_t)((((nc + j) >> 2) << 10) + (((nc + j) & 3) << 8))) = make_float2(v0, v1);
                        if (p.out_q) *reinterpret_cast<__nv_bfloat162 *>(p.out_q + (size_t)grow[h] * p.Cout + c) = __floats2bfloat162_rn(v0, v1);
                        if (p.out_a) {
                            const float a0f = live[h] ? fmaxf(fmaf(sa.x, v0, ta.x), 0.0f) : 0.0f, a1f = live[h] ? fmaxf(fmaf(sa.y, v1, ta.y), 0.0f) : 0.0f;
                            *reinterpret_cast<__nv_bfloat162 *>(p.out_a + (size_t)grow[h] * p.Cout + c) = __floats2bfloat162_rn(a0f, a1f);
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------- wide head convolution ----
struct HeadWideArgs {
    const int32_t *count;
    int max_count, H, W, Cin, K, P_pad, Wp, G;   // G boards per CTA iteration, G * P_pad a multiple of 16
    long long in_rows;        // allocated rows of the input
    const float *in;          // blocked fp32 padded rows [row][Cin]
    const uint4 *frags;       // [tap][k-step][n-tile (2)][lane] {hi0, hi1, lo0, lo1}: columns 0..7 head 1, 8..15 head 2
    const float *bias;        // [16]
    int cout1, cout2;         // live output channels of the two heads (<= 8 each; cout2 = 0: single head)
    float *out1, *out2;       // flat [leaf][H*W*cout] (H, W, C order)
};

// shared memory: fragments | hi rows | lo rows; a row is Cin bf16 = Cin/8 16-byte chunks, chunk index XOR (row & 7)
static inline size_t head_wide_smem(int Cin, int K, int P_pad, int Wp, int G) {
    const int halo = (K >> 1) * (Wp + 1), rows = G * P_pad + 2 * halo;
    return (size_t)K * K * (Cin >> 4) * 2 * 32 * 16 + (size_t)2 * rows * Cin * 2 + 16;
}

template <int K> __global__ void __launch_bounds__(512) headconv_wide_kernel(HeadWideArgs p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int taps = K * K, kh = K >> 1;
    const int KS = p.Cin >> 4, CH = p.Cin >> 3;
    const int halo = kh * (p.Wp + 1), srows = p.G * p.P_pad + 2 * halo;
    uint4 *s_frag = reinterpret_cast<uint4 *>(smem);
    uint8_t *s_hi = smem + (size_t)taps * KS * 2 * 32 * 16, *s_lo = s_hi + (size_t)srows * p.Cin * 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (int)(blockDim.x >> 5);
    for (int i = threadIdx.x; i < taps * KS * 2 * 32; i += blockDim.x) s_frag[i] = p.frags[i];
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const int n_groups = (cnt + p.G - 1) / p.G;
    const int n_mt = p.G * p.P_pad / 16, ncell = p.H * p.W;
    const uint32_t hi_base = (uint32_t)__cvta_generic_to_shared(s_hi), lo_base = (uint32_t)__cvta_generic_to_shared(s_lo);
    const int row_bytes = p.Cin * 2;
    // ldmatrix.x4: lane -> matrix (lane >> 3): rows (0-7 | 8-15) x k (0-7 | 8-15); row inside the matrix = lane & 7
    const int a_row = ((lane >> 3) & 1) * 8 + (lane & 7), a_chunk = lane >> 4;
    const int g = lane >> 2, t = lane & 3;
    const float b00 = p.bias[2 * t], b01 = p.bias[2 * t + 1], b10 = p.bias[8 + 2 * t], b11 = p.bias[8 + 2 * t + 1];
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        __syncthreads();
        // stage rows [row0 - halo, row0 + G*P_pad + halo) as bf16 hi / lo; padding positions, rows of dead boards and rows
        // outside the tensor become zeros (the fp32 stream may hold don't-care values there)
        const long long row0 = (long long)(grp * p.G) * p.P_pad - halo;
        for (int i0 = threadIdx.x; i0 < srows * CH; i0 += 4 * (int)blockDim.x) {     // four loads in flight per thread
            float v[4][8];
            bool ok[4];
            int pcs[4], srs[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * (int)blockDim.x;
                const int pc = i / srows, sr = i - pc * srows;       // consecutive threads: consecutive rows of one 8-channel piece
                pcs[u] = pc; srs[u] = sr;
                // board and position of the staged row in 32-bit arithmetic (halo < P_pad: at most one board before / after)
                const int rel = sr - halo;
                int b = grp * p.G, pos = rel;
                if (rel < 0) { b -= 1; pos = rel + p.P_pad; }
                else if (rel >= p.G * p.P_pad) { b += p.G; pos = rel - p.G * p.P_pad; }
                else { const int q = rel / p.P_pad; b += q; pos = rel - q * p.P_pad; }
                const int yy = pos / p.Wp;
                ok[u] = i < srows * CH && b >= 0 && b < cnt && yy != 0 && yy <= p.H && pos - yy * p.Wp != p.Wp - 1;
                if (ok[u]) gaz_conv::ldg256(p.in + f32_blk_index(row0 + sr, pc * 8, p.Cin), v[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (i0 + u * (int)blockDim.x >= srows * CH) continue;
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float x0 = ok[u] ? v[u][2 * e] : 0.0f, x1 = ok[u] ? v[u][2 * e + 1] : 0.0f;
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                    hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                const size_t off = (size_t)srs[u] * row_bytes + (size_t)(((pcs[u] & ~7) | ((pcs[u] ^ srs[u]) & 7)) << 4);
                *reinterpret_cast<uint4 *>(s_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4 *>(s_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
        __syncthreads();
        for (int mt = warp; mt < n_mt; mt += nwarp) {
            // one accumulator per (head, term): six independent MMA chains instead of two chains of three dependent MMAs
            float ahh[2][4], alh[2][4], ahl[2][4];
#pragma unroll
            for (int hd = 0; hd < 2; hd++)
#pragma unroll
                for (int e = 0; e < 4; e++) { ahh[hd][e] = 0.0f; alh[hd][e] = 0.0f; ahl[hd][e] = 0.0f; }
            for (int tp = 0; tp < taps; tp++) {
                const int sr = halo + mt * 16 + a_row + (tp / K - kh) * p.Wp + (tp % K - kh);
                const uint32_t rowoff = (uint32_t)(sr * row_bytes);
                const uint4 *fr = s_frag + (size_t)tp * KS * 64 + lane;
#pragma unroll 4
                for (int ks = 0; ks < KS; ks++) {
                    const int pc = ks * 2 + a_chunk;
                    const uint32_t coff = rowoff + (uint32_t)(((pc & ~7) | ((pc ^ sr) & 7)) << 4);
                    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(hi_base + coff));
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(l0), "=r"(l1), "=r"(l2), "=r"(l3) : "r"(lo_base + coff));
                    const uint4 f0 = fr[ks * 64], f1 = fr[ks * 64 + 32];
                    mma_bf16_16816(ahh[0], h0, h1, h2, h3, f0.x, f0.y);     // hi * hi
                    mma_bf16_16816(alh[0], l0, l1, l2, l3, f0.x, f0.y);     // lo * hi
                    mma_bf16_16816(ahl[0], h0, h1, h2, h3, f0.z, f0.w);     // hi * lo
                    if (p.cout2 > 0) {
                        mma_bf16_16816(ahh[1], h0, h1, h2, h3, f1.x, f1.y);
                        mma_bf16_16816(alh[1], l0, l1, l2, l3, f1.x, f1.y);
                        mma_bf16_16816(ahl[1], h0, h1, h2, h3, f1.z, f1.w);
                    }
                }
            }
            float acc[2][4];
#pragma unroll
            for (int hd = 0; hd < 2; hd++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[hd][e] = ahh[hd][e] + (alh[hd][e] + ahl[hd][e]);
            // C fragment: acc[.][0..1] = row g, columns 2t, 2t+1; acc[.][2..3] = row g + 8
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int rr = mt * 16 + g + 8 * h;
                const int gb = rr / p.P_pad, pos = rr - gb * p.P_pad;
                const int b = grp * p.G + gb;
                const int yy = pos / p.Wp - 1, xx = pos % p.Wp;
                if (b >= cnt || yy < 0 || yy >= p.H || xx >= p.W) continue;
                const int cell = yy * p.W + xx, co = 2 * t;
                if (co < p.cout1) {
                    float *o = p.out1 + ((size_t)b * ncell + cell) * p.cout1 + co;
                    o[0] = acc[0][2 * h] + b00;
                    if (co + 1 < p.cout1) o[1] = acc[0][2 * h + 1] + b01;
                }
                if (co < p.cout2) {
                    float *o = p.out2 + ((size_t)b * ncell + cell) * p.cout2 + co;
                    o[0] = acc[1][2 * h] + b10;
                    if (co + 1 < p.cout2) o[1] = acc[1][2 * h + 1] + b11;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- dense chain ----
struct MlpLayer {
    int In, Out, pre_affine, pre_relu, act;     // act: GAZ_ACT_* applied to this layer's output
    const float *w, *bias, *pre_scale, *pre_shift;   // w [In][Out]
};
struct MlpArgs {
    const int32_t *count;
    int max_count, n_layers;
    const uint4 *frag0;   // optional: layer 0 on mma.sync - its weights as B fragments [k-step][n-tile][lane] {hi0, hi1, lo0, lo1}
                          // (host_frag); needs In % 16 == 0; n-tiles padded to a multiple of 8 with zero fragments
    MlpLayer L[3];
    const float *in;      // [leaf][L[0].In]
    float *out;           // [leaf][L[n-1].Out]
};
// leaves per CTA: template parameter TL of the kernel (16: small batches, more CTAs; 32: twice the register tile per shared-memory
// load once there are enough CTAs to fill the SMs - 117 -> 81 us for the two Connect4 stacks at 4096 leaves)
constexpr int MLP_KT = 16;       // weight rows per shared-memory tile
// activations [feature][leaf] with a leading stride of TL + 4 floats: the A-fragment gathers of the mma.sync layer (address
// (k0 + 2t) * stride + g) then touch 32 different banks, the transposing stores of the load phase 8 instead of 1
static inline size_t mlp_smem(const MlpArgs &a, int MLP_TL) {
    if (a.frag0 && a.n_layers == 1) return (size_t)a.L[0].In * (MLP_TL + 4) * 4 + 16;   // only the input lives in shared memory
    int mx = 0, mo = 0;
    for (int i = 0; i < a.n_layers; i++) {
        mx = a.L[i].In > mx ? a.L[i].In : mx; mx = a.L[i].Out > mx ? a.L[i].Out : mx;
        mo = a.L[i].Out > mo ? a.L[i].Out : mo;
    }
    return (size_t)(2 * mx * (MLP_TL + 4) + MLP_KT * mo) * 4 + 16;
}

// One CTA = MLP_TL leaves through every layer of the chain.  Activations live in shared memory as [feature][leaf]; the
// weights of a layer stream through a [MLP_KT][Out] shared-memory tile (the tile is a contiguous piece of the [In][Out]
// matrix: coalesced loads, the next tile is fetched into registers while the current one is multiplied); a thread owns 4
// consecutive outputs x `lpg` leaves.  Terms are accumulated in k order.  Requires Out <= 128 (8 prefetch registers).
template <int LPG, int MLP_TL>
__device__ __forceinline__ void mlp_tile(float (&acc)[8][4], const float *s_w, const float *ap0, int kn, int Out, int o0, bool vec) {
    for (int kk = 0; kk < kn; kk++) {
        float w4[4];
        if (vec) {
            const float4 ww = *reinterpret_cast<const float4 *>(s_w + kk * Out + o0);
            w4[0] = ww.x; w4[1] = ww.y; w4[2] = ww.z; w4[3] = ww.w;
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++) w4[b] = o0 + b < Out ? s_w[kk * Out + o0 + b] : 0.0f;
        }
        const float *ap = ap0 + kk * (MLP_TL + 4);
        float av[LPG];
        if (LPG >= 4) {
#pragma unroll
            for (int a = 0; a < LPG; a += 4) {
                const float4 t4 = *reinterpret_cast<const float4 *>(ap + a);
                av[a] = t4.x; av[a + 1] = t4.y; av[a + 2] = t4.z; av[a + 3] = t4.w;
            }
        } else {
#pragma unroll
            for (int a = 0; a < LPG; a++) av[a] = ap[a];
        }
#pragma unroll
        for (int a = 0; a < LPG; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = fmaf(av[a], w4[b], acc[a][b]);
    }
}

// Two chains with independent inputs (the policy and the value stack of Connect4 / TicTacToe) share one launch: blockIdx.y
// selects the chain, so the second stack runs in the shadow of the first instead of as another latency-bound launch.
template <int MLP_TL> __global__ void __launch_bounds__(256) mlp_chain_kernel(const MlpArgs p0, const MlpArgs p1) {
    extern __shared__ __align__(16) float s_act[];
    const MlpArgs &p = blockIdx.y ? p1 : p0;
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const int leaf0 = blockIdx.x * MLP_TL;
    if (leaf0 >= cnt) return;
    int mx = 0;
    for (int i = 0; i < p.n_layers; i++) { mx = max(mx, max(p.L[i].In, p.L[i].Out)); }
    constexpr int LS = MLP_TL + 4;   // leading stride of the [feature][leaf] activations
    float *bufA = s_act, *bufB = s_act + (size_t)mx * LS, *s_w = s_act + (size_t)2 * mx * LS;
    {   // load + pre-activation of the first layer's input: [leaf][In] -> [In][leaf]; four independent loads in flight per thread
        const MlpLayer &l = p.L[0];
        const int tot = l.In * MLP_TL;
        for (int i0 = threadIdx.x; i0 < tot; i0 += 4 * (int)blockDim.x) {
            float a[4], sc[4], sh[4];
            int kk[4], lf[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * (int)blockDim.x;
                lf[u] = i / l.In; kk[u] = i - lf[u] * l.In;
                const bool ok = i < tot && leaf0 + lf[u] < cnt;
                a[u] = ok ? p.in[(size_t)(leaf0 + lf[u]) * l.In + kk[u]] : 0.0f;
                sc[u] = (ok && l.pre_affine) ? l.pre_scale[kk[u]] : 1.0f;
                sh[u] = (ok && l.pre_affine) ? l.pre_shift[kk[u]] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (i0 + u * (int)blockDim.x >= tot) continue;
                float x = l.pre_affine ? fmaf(sc[u], a[u], sh[u]) : a[u];
                if (l.pre_relu) x = fmaxf(x, 0.0f);
                if (leaf0 + lf[u] >= cnt) x = 0.0f;
                bufA[kk[u] * LS + lf[u]] = x;
            }
        }
    }
    __syncthreads();
    int li0 = 0;
    if (p.frag0) {
        // ---- layer 0 on the warp-level tensor cores (Connect4: 336 -> 128 is 86 % of a stack's FLOPs and the CUDA-core form is
        // bound by its shared-memory loads).  Activations and weights are split into bf16 hi + lo parts, three MMAs per fragment
        // (hi*hi + lo*hi + hi*lo) as in headconv_wide_kernel: fp32-level accuracy.  Warp w: m-tile w % MT (16 leaves), n-tiles
        // (w / MT) * NPW .. + NPW; A fragments are gathered from the [feature][leaf] activations and split on the fly.
        const MlpLayer &l = p.L[0];
        constexpr int MT = MLP_TL / 16;
        const int NT = (((l.Out + 7) >> 3) + 7) & ~7, NPW = NT / (8 / MT);   // n-tiles (padded to a multiple of 8; the fragments beyond Out
                                                                             // are zero) and n-tiles per warp: <= 8 for Out <= 256 (TL 32)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
        const int mt = warp % MT, nt0 = (warp / MT) * NPW;
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f; }
        const float *ab = bufA + mt * 16 + g;
        const int KS = l.In >> 4;
        auto split = [](float x0, float x1, uint32_t &hi, uint32_t &lo) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
            hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        };
#pragma unroll 1
        for (int ks = 0; ks < KS; ks++) {
            const float *a = ab + (size_t)(ks * 16 + 2 * t) * LS;
            uint32_t h[4], lo[4];
            split(a[0], a[LS], h[0], lo[0]);                      // row g,     k = 2t, 2t+1
            split(a[8], a[LS + 8], h[1], lo[1]);                  // row g + 8
            split(a[8 * LS], a[9 * LS], h[2], lo[2]);             // row g,     k = 2t+8, 2t+9
            split(a[8 * LS + 8], a[9 * LS + 8], h[3], lo[3]);     // row g + 8
            const uint4 *fr = p.frag0 + ((size_t)ks * NT + nt0) * 32 + lane;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j < NPW) {
                    const uint4 f = __ldg(fr + j * 32);
                    mma_bf16_16816(acc[j], h[0], h[1], h[2], h[3], f.x, f.y);       // hi * hi
                    mma_bf16_16816(acc[j], lo[0], lo[1], lo[2], lo[3], f.x, f.y);   // lo * hi
                    mma_bf16_16816(acc[j], h[0], h[1], h[2], h[3], f.z, f.w);       // hi * lo
                }
            }
        }
        const bool last = p.n_layers == 1;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j >= NPW) continue;
#pragma unroll
            for (int e = 0; e < 4; e++) {     // C fragment: e 0,1 = row g, columns 2t, 2t+1; e 2,3 = row g + 8
                const int lf = mt * 16 + g + (e >> 1) * 8, o = (nt0 + j) * 8 + 2 * t + (e & 1);
                if (o >= l.Out) continue;
                float v = acc[j][e] + l.bias[o];
                if (l.act == 1) v = fmaxf(v, 0.0f);
                else if (l.act == 3) v = tanhf(v);
                if (last) {
                    if (leaf0 + lf < cnt) p.out[(size_t)(leaf0 + lf) * l.Out + o] = v;
                } else {
                    const MlpLayer &nx = p.L[1];
                    if (nx.pre_affine) v = fmaf(nx.pre_scale[o], v, nx.pre_shift[o]);
                    if (nx.pre_relu) v = fmaxf(v, 0.0f);
                    bufB[o * LS + lf] = v;
                }
            }
        }
        __syncthreads();
        float *tswap = bufA; bufA = bufB; bufB = tswap;
        li0 = 1;
    }
    for (int li = li0; li < p.n_layers; li++) {
        const MlpLayer &l = p.L[li];
        const bool last = li + 1 == p.n_layers;
        const int n_o4 = (l.Out + 3) >> 2;                  // threads per leaf group
        const int groups = min((int)blockDim.x / n_o4, MLP_TL);
        int lpg = (MLP_TL + groups - 1) / groups;           // leaves per group, rounded up to 1 / 2 / 4 / 8 (Out <= 128)
        lpg = lpg <= 1 ? 1 : (lpg <= 2 ? 2 : (lpg <= 4 ? 4 : 8));
        const int og = threadIdx.x % n_o4, grp = threadIdx.x / n_o4;
        const int o0 = og * 4, lf0 = grp * lpg;
        const bool active = grp < groups && lf0 < MLP_TL;
        const bool vec = (l.Out & 3) == 0;
        const int tile = MLP_KT * l.Out, total = l.In * l.Out;
        float acc[8][4];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.0f;
        float pre[8];
        auto fetch = [&](int k0) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = (int)threadIdx.x + u * 256, gi = k0 * l.Out + e;
                pre[u] = (e < tile && gi < total) ? __ldg(l.w + gi) : 0.0f;
            }
        };
        fetch(0);
        for (int k0 = 0; k0 < l.In; k0 += MLP_KT) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int e = (int)threadIdx.x + u * 256;
                if (e < tile) s_w[e] = pre[u];
            }
            __syncthreads();
            if (k0 + MLP_KT < l.In) fetch(k0 + MLP_KT);
            if (active) {
                const int kn = min(MLP_KT, l.In - k0);
                const float *ap0 = bufA + k0 * LS + lf0;
                switch (lpg) {     // static trip counts: the generic predicated form spent 6x the useful instructions
                case 1: mlp_tile<1, MLP_TL>(acc, s_w, ap0, kn, l.Out, o0, vec); break;
                case 2: mlp_tile<2, MLP_TL>(acc, s_w, ap0, kn, l.Out, o0, vec); break;
                case 4: mlp_tile<4, MLP_TL>(acc, s_w, ap0, kn, l.Out, o0, vec); break;
                default: mlp_tile<8, MLP_TL>(acc, s_w, ap0, kn, l.Out, o0, vec); break;
                }
            }
            __syncthreads();
        }
        if (active) {
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const int lf = lf0 + a;
                if (a >= lpg || lf >= MLP_TL) continue;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int o = o0 + b;
                    if (o >= l.Out) continue;
                    float v = acc[a][b] + l.bias[o];
                    if (l.act == 1) v = fmaxf(v, 0.0f);
                    else if (l.act == 3) v = tanhf(v);
                    if (last) {
                        if (leaf0 + lf < cnt) p.out[(size_t)(leaf0 + lf) * l.Out + o] = v;
                    } else {
                        const MlpLayer &nx = p.L[li + 1];
                        if (nx.pre_affine) v = fmaf(nx.pre_scale[o], v, nx.pre_shift[o]);
                        if (nx.pre_relu) v = fmaxf(v, 0.0f);
                        bufB[o * LS + lf] = v;
                    }
                }
            }
        }
        __syncthreads();
        float *tswap = bufA; bufA = bufB; bufB = tswap;
    }
}

} // namespace gaz_small
