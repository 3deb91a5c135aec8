// gaz_convt.cuh -- channel-on-lanes ("transposed") implicit-GEMM convolution on tcgen05: EXPERIMENTAL v3 of the trunk
// convolution, enabled with GAZ_CONV_T=1 (parity-tested, not the default).  Measured on B200 at 16384 Gomoku leaves:
// plain 3x3 C128 convolution 0.75 ms (v2 pair: 0.75 ms), fused-SE convolution 1.8-2.1 ms (v2 pair: 1.63 ms): the
// cheaper per-channel epilogue arithmetic did not pay because the fused-SE epilogue is bound by its serial phases and
// HBM traffic, not by instruction count - so v2 stays the product path and this file documents the experiment.
//
// v2 (gaz_conv.cuh) computes D[rows x cout] = X[rows x K] * W^T, so an epilogue thread owns one board ROW and all
// cout channels: every per-channel parameter is a broadcast load, the SE board means need a 31-shuffle butterfly per
// 32 values, and ncu showed the fused-SE epilogue issue/latency bound at ~40 SASS instructions per output element
// (profiles/r01_conv_v2_*).  v3 swaps the operand roles:
//
//      D^T[cout x 256 positions] = W[cout x K] (A operand, 128 x 64 weight tile, K-major)
//                                 * X^T        (B operand = the tile's activation slab, N = 256 rows, K-major,
//                                               filter taps = row-shifted descriptors exactly as in v2)
//
// so the accumulator has CHANNELS on the TMEM lanes and board positions on the columns.  An epilogue thread now owns
// one channel: bias / BN scale / BN shift / SE gate are scalars in registers, the SE board mean is a plain per-thread
// sum over columns, and the padding mask is warp-uniform.  One MMA (M = 128, N = 256, K = 16) covers a whole 256-row
// tile, 72 MMAs per 3x3 C128 layer tile; shared-memory operand traffic is 12 KB per 128-cycle MMA (96 B/clk) without
// needing a CTA pair.
// Warp roles (640 threads): 0 slab TMA producer, 1 weight TMA producer, 2 MMA issuer (+TMEM alloc), 3 idle,
// 4..19 epilogue in two groups of 8 warps that take alternate tiles (group g = accumulator stage g): TMEM lane
// quarter q = warp & 3 (channels 32q..32q+31), position half (128 columns) from the warp index.
// fp32 row tensors (residual stream) use an 8-row interleaved layout, f32_t_index: element (row, c) lives at
// ((row >> 3) * C + c) * 8 + (row & 7), so a warp (32 consecutive channels) reading 8 positions touches 1 KB
// contiguous with one 256-bit vector per lane.  bf16 operands of the next layer stay row-major [row][C] (TMA needs
// that) and are written with 16-bit stores, one position per instruction = 64 contiguous bytes per warp.
#pragma once
#include "gaz_conv.cuh"

namespace gaz_convt {
using namespace gaz_tc;
using gaz_conv::BoardConvArgs;
using gaz_conv::HALO;
using gaz_conv::Ring;
using gaz_conv::SLAB_BOX_ROWS;
using gaz_conv::SLAB_BYTES;
using gaz_conv::TILE_ROWS;
using gaz_conv::ldg256;
using gaz_conv::named_bar_sync;
using gaz_conv::stg256;

__host__ __device__ __forceinline__ size_t f32_t_index(long long row, int c, int C) {
    return ((size_t)((row >> 3) * C + c) << 3) + (size_t)(row & 7);
}

struct TCfg {
    static constexpr int NSLAB = 4;
    static constexpr int NW = 4;                 // weight-tile ring
    static constexpr int W_BYTES = 128 * 128;    // 128 output channels x 64 bf16
    static constexpr int STAGE_BYTES = 0;
    static constexpr int TMEM_COLS = 512;        // 2 accumulator stages x 256 positions
    static constexpr int SE_FLOATS = 2 * (2 * 128 + 4 * 64 + 2 * 128); // per epilogue group: sums [2][128], hidden partials [4][64], gate partials [2][128]
    static constexpr int SMEM = NSLAB * SLAB_BYTES + NW * W_BYTES + STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                                SE_FLOATS * 4;
};

// kind::f16 instruction descriptor for M = 128, N = 256 (bf16 x bf16 -> f32, both operands K-major)
__host__ __device__ constexpr uint32_t idesc_m128_n256() { return umma_idesc_bf16(128, 256); }

__global__ void __launch_bounds__(640, 1)
conv_t_kernel(const __grid_constant__ CUtensorMap tmA /*activations, 152-row box*/,
              const __grid_constant__ CUtensorMap tmW /*weights, 128-row box (rows beyond cout zero-filled)*/,
              const __grid_constant__ CUtensorMap tmOa, const __grid_constant__ CUtensorMap tmOb /*32ch x 16row boxes*/,
              const __grid_constant__ BoardConvArgs p, const int cout) {
    using Cfg = TCfg;
    constexpr int NSLAB = Cfg::NSLAB, NW = Cfg::NW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;
    uint8_t *sW = base + NSLAB * SLAB_BYTES;
    uint8_t *sStage = sW + NW * Cfg::W_BYTES;
    uint64_t *bars = (uint64_t *)(sStage + Cfg::STAGE_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + NSLAB, *w_full = bars + 2 * NSLAB, *w_empty = bars + 2 * NSLAB + NW;
    uint64_t *tfull = bars + 2 * NSLAB + 2 * NW, *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);
    float *s_se = (float *)(bars + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const long long valid_rows = (long long)cnt * p.P_pad;
    const int n_tiles = (int)((valid_rows + TILE_ROWS - 1) / TILE_ROWS);
    const int dbg = p.base_offset_mode; // timing experiments: 4 no stores, 8 no residual loads, 64 no SE passes

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLAB; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NW; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        if (p.out_a) tma_prefetch_desc(&tmOa);
        if (p.out_b) tma_prefetch_desc(&tmOb);
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) { // ---------------- activation-slab TMA producer
            Ring r;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int row0 = t * TILE_ROWS - HALO;
                for (int kc = 0; kc < p.kpt; kc++) {
                    mbar_wait(&a_empty[r.idx], r.phase ^ 1);
                    uint8_t *dst = sA + r.idx * SLAB_BYTES;
                    mbar_expect_tx(&a_full[r.idx], SLAB_BYTES);
                    tma_load_2d(dst, &tmA, &a_full[r.idx], kc * 64, row0);
                    tma_load_2d(dst + SLAB_BYTES / 2, &tmA, &a_full[r.idx], kc * 64, row0 + SLAB_BOX_ROWS);
                    r.advance(NSLAB);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) { // ---------------- weight-tile TMA producer: order (K-block, tap)
            Ring r;
            const int cin = p.kpt * 64;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
                for (int kc = 0; kc < p.kpt; kc++)
                    for (int tap = 0; tap < p.taps; tap++) {
                        mbar_wait(&w_empty[r.idx], r.phase ^ 1);
                        mbar_expect_tx(&w_full[r.idx], (uint32_t)(cout < 128 ? cout : 128) * 128u); // box rows = min(cout, 128)
                        tma_load_2d(sW + r.idx * Cfg::W_BYTES, &tmW, &w_full[r.idx], tap * cin + kc * 64, 0);
                        r.advance(NW);
                    }
        }
    } else if (warp == 2) { // ---------------- MMA issuer: whole warp, uniform control flow, one elected lane per instruction
        constexpr uint32_t idesc = idesc_m128_n256();
        Ring ra, rw;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)(acc * TILE_ROWS);
            for (int kc = 0; kc < p.kpt; kc++) {
                mbar_wait(&a_full[ra.idx], ra.phase);
                const uint32_t slab_lo = umma_desc_lo(smem_u32(sA + ra.idx * SLAB_BYTES) + (uint32_t)(HALO * 128));
                int dy = -1, dx = -1; // tap = (dy + 1) * 3 + (dx + 1)
                for (int tap = 0; tap < p.taps; tap++) {
                    mbar_wait(&w_full[rw.idx], rw.phase);
                    tc_fence_after();
                    const int shift = p.taps == 9 ? dy * p.Wp + dx : 0;
                    const uint32_t w_lo = umma_desc_lo(smem_u32(sW + rw.idx * Cfg::W_BYTES));
                    const uint32_t x_lo = slab_lo + (uint32_t)(shift * 8); // 128 B per row = 8 descriptor units
                    const uint32_t accf = (uint32_t)((kc | tap) != 0);     // the first MMA of a tile overwrites
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        umma_bf16_elect<false>(d0, w_lo + (uint32_t)(k * 2), x_lo + (uint32_t)(k * 2), idesc, k == 0 ? accf : 1u);
                    umma_commit_elect<false>(&w_empty[rw.idx]);
                    rw.advance(NW);
                    if (++dx == 2) { dx = -1; dy++; }
                }
                umma_commit_elect<false>(&a_empty[ra.idx]);
                ra.advance(NSLAB);
            }
            umma_commit_elect<false>(&tfull[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) { // ---------------- epilogue: two groups of 8 warps, group g owns accumulator stage g
        // Tiles alternate between the groups, so the serial parts of one tile's epilogue (SE barriers, the tiny
        // dense layers' L2 latency, residual loads) overlap with the other group's tile instead of stalling the SM.
        const int ew = warp - 4;
        const int grp = ew >> 3;                    // 0 / 1 = accumulator stage
        const int q = warp & 3, pg = (ew >> 2) & 1; // channel quarter, position half (128 columns)
        const int c = q * 32 + lane;                // this thread's output channel
        const bool chan_ok = c < cout;
        const int gt = threadIdx.x - 128 - grp * 256; // 0..255 inside the group
        float *s_grp = s_se + grp * (Cfg::SE_FLOATS / 2);
        float *s_part = s_grp, *s_hp = s_grp + 2 * 128, *s_gp = s_hp + 4 * 64; // [2][128] sums, [4][64] hidden, [2][128] gate partials
        // per-channel parameters: registers for the whole kernel (global loads: a lane-indexed LDC would serialise)
        const float bias = __ldg(p.d_par + c), sc_a = __ldg(p.d_par + 128 + c), sh_a = __ldg(p.d_par + 256 + c),
                    sc_b = __ldg(p.d_par + 384 + c), sh_b = __ldg(p.d_par + 512 + c);
        const bool plain = !p.res && !p.out_raw && !p.se;           // conv1: only relu(BN(conv + bias)) leaves
        const float shf_a = fmaf(sc_a, bias, sh_a), shf_b = fmaf(sc_b, bias, sh_b);
        const float se_b2c = (p.se && chan_ok) ? __ldg(p.se_b2 + c) : 0.0f;
        const float inv_cells = p.se ? 1.0f / (float)p.n_cells : 0.0f;
        const int acc = grp;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x + grp * gridDim.x; t < n_tiles; t += 2 * gridDim.x) {
            const long long row_base = (long long)t * TILE_ROWS + pg * 128; // first of this warp's 128 positions
            // padding masks of the warp's 128 positions, warp-uniform: bit j of m[k] = position row_base + 32k + j is a live cell
            uint32_t m[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long long r0 = row_base + k * 32 + lane;
                const int p0 = (int)(r0 % p.P_pad);
                m[k] = __ballot_sync(0xffffffffu, r0 < valid_rows && (p0 / p.Wp) != 0 && (p0 % p.Wp) != p.Wp - 1);
            }
            const bool use_res = p.res && !(dbg & 8) && chan_ok;
            float rnext[16]; // residual of the first chunk: issued before the accumulator is even ready
            if (use_res) {
                const float *rp = p.res + f32_t_index(row_base, c, cout);
#pragma unroll 1
                for (int k = 2; k < 16; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)k * cout * 8)); // rest of the tile -> L2
                ldg256(rp, *reinterpret_cast<float(*)[8]>(&rnext[0]));
                ldg256(rp + (size_t)cout * 8, *reinterpret_cast<float(*)[8]>(&rnext[8]));
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TILE_ROWS + pg * 128);
            float gate = 1.0f;
            if (p.se && !(dbg & 64)) {
                // ---- pass 1: this channel's sum over the warp's live positions (tile == board)
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int ck = 0; ck < 8; ck += 2) {
                    uint32_t r0[16], r1[16];
                    tmem_ld_32x16(t_acc + (uint32_t)(ck * 16), r0);
                    tmem_ld_32x16(t_acc + (uint32_t)(ck * 16 + 16), r1);
                    tmem_ld_wait();
                    const uint32_t mk = m[ck >> 1];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        if ((mk >> j) & 1u) s0 += __uint_as_float(r0[j]);          // warp-uniform predicates
                        if ((mk >> (16 + j)) & 1u) s1 += __uint_as_float(r1[j]);
                    }
                }
                s_part[pg * 128 + c] = s0 + s1;
                named_bar_sync(1 + grp, 256);
                {   // dense1 (C -> R) over 256 threads: output j, quarter `part` of the inputs; the conv bias is folded
                    // into the dense bias on the host (se_b1 already holds b1 + W1^T bias)
                    const int j = gt & 63, part = gt >> 6;
                    if (j < p.se_r) {
                        const float *w1 = p.se_w1 + (size_t)(part * 32) * p.se_r + j;
                        float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const int ii = part * 32 + i;
                            h0 = fmaf((s_part[ii] + s_part[128 + ii]) * inv_cells, __ldg(w1 + (size_t)(i) * p.se_r), h0);
                            h1 = fmaf((s_part[ii + 1] + s_part[129 + ii]) * inv_cells, __ldg(w1 + (size_t)(i + 1) * p.se_r), h1);
                            h2 = fmaf((s_part[ii + 2] + s_part[130 + ii]) * inv_cells, __ldg(w1 + (size_t)(i + 2) * p.se_r), h2);
                            h3 = fmaf((s_part[ii + 3] + s_part[131 + ii]) * inv_cells, __ldg(w1 + (size_t)(i + 3) * p.se_r), h3);
                        }
                        s_hp[part * 64 + j] = (h0 + h1) + (h2 + h3);
                    }
                }
                named_bar_sync(1 + grp, 256);
                {   // dense2 (R -> C): output channel cc, half `part` of the hidden units (relu(dense1) rebuilt on the fly)
                    const int cc = gt & 127, part = gt >> 7;
                    const int r2 = p.se_r >> 1;
                    const float *w2 = p.se_w2 + (size_t)(part * r2) * 128 + cc;
                    float g0 = 0.0f, g1 = 0.0f;
#pragma unroll 8
                    for (int i = 0; i < r2; i += 2) {
                        const int ii = part * r2 + i;
                        const float ha = fmaxf(((s_hp[ii] + s_hp[64 + ii]) + (s_hp[128 + ii] + s_hp[192 + ii])) + __ldg(p.se_b1 + ii), 0.0f);
                        const float hb = fmaxf(((s_hp[ii + 1] + s_hp[65 + ii]) + (s_hp[129 + ii] + s_hp[193 + ii])) + __ldg(p.se_b1 + ii + 1), 0.0f);
                        g0 = fmaf(ha, __ldg(w2 + (size_t)(i) * 128), g0);
                        g1 = fmaf(hb, __ldg(w2 + (size_t)(i + 1) * 128), g1);
                    }
                    s_gp[part * 128 + cc] = g0 + g1;
                }
                named_bar_sync(1 + grp, 256);
                gate = 1.0f / (1.0f + expf(-((s_gp[c] + s_gp[128 + c]) + se_b2c)));
            }
            const float bg = bias * gate; // gate * (acc + bias) + res = fma(acc, gate, bg) + res
            // ---- output pass over the warp's eight 16-position chunks
#pragma unroll 1
            for (int ck = 0; ck < 8; ck++) {
                const long long row0 = row_base + ck * 16;
                const uint32_t mk = (m[ck >> 1] >> ((ck & 1) * 16)) & 0xffffu;
                uint32_t r[16];
                tmem_ld_32x16(t_acc + (uint32_t)(ck * 16), r);
                float rcur[16];
                if (use_res) {
#pragma unroll
                    for (int j = 0; j < 16; j++) rcur[j] = rnext[j];
                    if (ck + 1 < 8) {
                        const float *rp = p.res + f32_t_index(row0 + 16, c, cout);
                        ldg256(rp, *reinterpret_cast<float(*)[8]>(&rnext[0]));
                        ldg256(rp + (size_t)cout * 8, *reinterpret_cast<float(*)[8]>(&rnext[8]));
                    }
                }
                tmem_ld_wait();
                if (dbg & 4) continue; // timing experiment only: no output stores
                float v[16];
                if (!plain) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float x = fmaf(__uint_as_float(r[j]), gate, bg);
                        if (use_res) x += rcur[j];
                        v[j] = ((mk >> j) & 1u) ? x : 0.0f;                   // padding positions stay exactly zero
                    }
                    if (p.out_raw && chan_ok) {
                        float *op = p.out_raw + f32_t_index(row0, c, cout);
                        stg256(op, *reinterpret_cast<float(*)[8]>(&v[0]));
                        stg256(op + (size_t)cout * 8, *reinterpret_cast<float(*)[8]>(&v[8]));
                    }
                }
                // bf16 operands of the next layer, row-major [row][cout]: lane = channel; lane pairs (2i, 2i+1) trade
                // values so that the even lane holds both channels of the even positions and the odd lane those of
                // the odd positions -> 32-bit stores, 64 contiguous bytes per position per warp (the row-per-lane
                // form of v2 touched 32 lines per store instruction)
#pragma unroll
                for (int o = 0; o < 2; o++) {
                    __nv_bfloat16 *outp = o == 0 ? p.out_a : p.out_b;
                    if (!outp || (dbg & 32)) continue;
                    const float sc = o == 0 ? sc_a : sc_b;
                    const float sh = plain ? (o == 0 ? shf_a : shf_b) : (o == 0 ? sh_a : sh_b);
                    const bool odd = lane & 1;
                    uint32_t *op = reinterpret_cast<uint32_t *>(outp + (size_t)(row0 + (odd ? 1 : 0)) * cout + (c & ~1));
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float x0 = plain ? __uint_as_float(r[j]) : v[j];
                        const float x1 = plain ? __uint_as_float(r[j + 1]) : v[j + 1];
                        float a0 = fmaxf(fmaf(sc, x0, sh), 0.0f), a1 = fmaxf(fmaf(sc, x1, sh), 0.0f);
                        a0 = ((mk >> j) & 1u) ? a0 : 0.0f;
                        a1 = ((mk >> (j + 1)) & 1u) ? a1 : 0.0f;
                        const float send = odd ? a0 : a1;                    // what the partner needs
                        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                        const float lo = odd ? recv : a0, hi = odd ? a1 : recv; // (channel c&~1, channel c|1) at my position
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                        if (chan_ok) op[(size_t)(j >> 1) * cout] = *reinterpret_cast<uint32_t *>(&h2); // rows j / j+1: 2*cout bf16 = cout words
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

} // namespace gaz_convt
