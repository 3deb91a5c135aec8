// gaz_block.cuh -- one pre-activation residual block (Net/ResNet/ResNet_Block.py:27-41 + Net/SE/SE_Block.py:15-23) as ONE
// kernel for tile == board geometries (Gomoku: 256 padded rows per board), C_in = C_out = 128:
//
//   a_x (bf16, HBM) --TMA--> slab X --conv1 MMAs--> TMEM acc1 --epilogue 1: relu(BN2(. + b1))--> slab H (bf16, SMEM only)
//        --conv2 MMAs--> TMEM acc2 --epilogue 2: SE gate * (. + b2) + residual--> x_out (fp32) and relu(BN(x_out)) (bf16)
//
// The bf16 intermediate between the two convolutions never leaves the SM (2.1 GB per block and launch at 16384 leaves),
// and conv1 - a tensor-bound kernel on its own - runs in the shadow of the memory-bound fused-SE epilogue of the
// previous board.  Every board depends only on itself: the rows a live output reads outside its own board are padding
// rows/columns (zeros), so slab H needs no data from neighbouring boards (its halo rows stay zero).
// CTA pairs (cta_group::2) as in gaz_conv.cuh: each CTA owns its board, slabs, TMEM and epilogues; the leader issues
// every MMA for both (M = 256) and each CTA stages half of every weight tile.
// Warp roles (512 threads): 0 slab-X TMA producer, 1 weight TMA producer (W1 tiles then W2 tiles per board), 2 MMA
// issuer, 3 idle, 4..7 epilogue 1 (one per TMEM lane quarter), 8..15 epilogue 2 (quarter x 128-row half).
// TMEM: two accumulator sets of 256 columns (2 x 128-row halves x 128 channels); boards alternate between the sets and
// BOTH convolutions of a board use the board's set (conv1 -> epilogue 1 drains it into slab H -> conv2 -> epilogue 2),
// so conv1, epilogue 1 and conv2 of board i+1 all run while epilogue 2 of board i is still reading the other set.
#pragma once
#include "gaz_conv.cuh"

namespace gaz_block {
using namespace gaz_tc;
using gaz_conv::f32_blk_index;
using gaz_conv::HALO;
using gaz_conv::ldg256;
using gaz_conv::named_bar_sync;
using gaz_conv::Ring;
using gaz_conv::SLAB_BOX_ROWS;
using gaz_conv::SLAB_BYTES;
using gaz_conv::stg256;
using gaz_conv::TILE_ROWS;

struct BlockArgs {
    const int32_t *count;
    int max_count;
    int Wp, n_cells, dbg;
    float par1[3 * 128];      // conv1 bias | BN2 scale | BN2 shift          (constant bank, uniform loads)
    float par2[5 * 128];      // conv2 bias | scale_a | shift_a | scale_b | shift_b
    const float *res;         // blocked fp32 residual stream in
    float *out_raw;           // blocked fp32 residual stream out
    __nv_bfloat16 *out_a, *out_b;
    int se, se_r;
    const float *se_w1, *se_b1, *se_w2, *se_b2; // se_b1 has the conv2 bias folded in (b1 + W1^T bias2)
};

struct Cfg {
    static constexpr int NW = 4;                    // weight-tile ring (half tiles: 64 output channels x 64 k)
    static constexpr int W_BYTES = 64 * 128;
    static constexpr int STAGE_BYTES = 8 * 2 * 2048; // epilogue-2 warps: one 32 x 32-channel bf16 tile per output
    static constexpr int SE_FLOATS = 8 * 128 + 3 * 128;
    static constexpr int SMEM = 4 * SLAB_BYTES + NW * W_BYTES + STAGE_BYTES + 1024 + 256 + SE_FLOATS * 4;
};

__global__ void __launch_bounds__(512, 1)
res_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOa,
                 const __grid_constant__ CUtensorMap tmOb, const __grid_constant__ BlockArgs p) {
    constexpr int BN = 128, NW = Cfg::NW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sX = base;                       // 2 slabs: K-blocks 0/1 of the block input
    uint8_t *sH = base + 2 * SLAB_BYTES;      // 2 slabs: K-blocks 0/1 of the conv1 output
    uint8_t *sW = base + 4 * SLAB_BYTES;
    uint8_t *sStage = sW + NW * Cfg::W_BYTES;
    uint64_t *bars = (uint64_t *)(sStage + Cfg::STAGE_BYTES);
    uint64_t *x_full = bars, *x_empty = bars + 2, *w_full = bars + 4, *w_empty = bars + 4 + NW;
    uint64_t *acc1_full = bars + 4 + 2 * NW, *e1_done = acc1_full + 1, *h_empty = acc1_full + 2,
             *acc2_full = acc1_full + 3 /*[2]*/, *acc2_empty = acc1_full + 5 /*[2]*/;
    uint32_t *tmem_slot = (uint32_t *)(acc1_full + 7);
    float *s_se = (float *)(bars + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int pair0 = (int)(blockIdx.x >> 1), pair_step = (int)(gridDim.x >> 1);
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const int n_tiles = cnt;                      // tile == board
    const int n_loop = (n_tiles + 1) / 2;
    const long long valid_rows = (long long)cnt * TILE_ROWS;
    const int dbg = p.dbg;

    // slab H halo rows are never written by epilogue 1: zero them once (rows [0, HALO) and [HALO + 256, 304) of both slabs)
    for (int i = threadIdx.x; i < 2 * 2 * HALO * 8; i += blockDim.x) {
        const int slab = i / (2 * HALO * 8), r = (i / 8) % (2 * HALO), ch = i & 7;
        const int row = r < HALO ? r : TILE_ROWS + r;
        *reinterpret_cast<uint4 *>(sH + slab * SLAB_BYTES + row * 128 + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
        for (int s = 0; s < NW; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(acc1_full, 1);
        mbar_init(e1_done, 8);      // 4 epilogue-1 warps of each CTA arrive on the leader's copy
        mbar_init(h_empty, 1);
        for (int a = 0; a < 2; a++) { mbar_init(&acc2_full[a], 1); mbar_init(&acc2_empty[a], 16); } // 8 epilogue-2 warps of each CTA
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        if (p.out_a) tma_prefetch_desc(&tmOa);
        if (p.out_b) tma_prefetch_desc(&tmOb);
    }
    fence_proxy_async(); // the zeroed halo rows must be visible to the MMA (async proxy) reads
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) { // ---------------- slab-X TMA producer
            uint32_t ph = 0;
            for (int lt = pair0; lt < n_loop; lt += pair_step, ph ^= 1) {
                const int t = 2 * lt + rank;
                const int row0 = t * TILE_ROWS - HALO;
                for (int kc = 0; kc < 2; kc++) {
                    mbar_wait(&x_empty[kc], ph ^ 1);
                    if (rank == 0) mbar_expect_tx(&x_full[kc], 2 * SLAB_BYTES);
                    uint8_t *dst = sX + kc * SLAB_BYTES;
                    tma_load_2d_pair(dst, &tmA, &x_full[kc], kc * 64, row0);
                    tma_load_2d_pair(dst + SLAB_BYTES / 2, &tmA, &x_full[kc], kc * 64, row0 + SLAB_BOX_ROWS);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) { // ---------------- weight-tile TMA producer: per board W1 (kc, tap) then W2 (kc, tap)
            Ring r;
            for (int lt = pair0; lt < n_loop; lt += pair_step)
                for (int cv = 0; cv < 2; cv++)
                    for (int kc = 0; kc < 2; kc++)
                        for (int tap = 0; tap < 9; tap++) {
                            mbar_wait(&w_empty[r.idx], r.phase ^ 1);
                            if (rank == 0) mbar_expect_tx(&w_full[r.idx], 2 * Cfg::W_BYTES);
                            tma_load_2d_pair(sW + r.idx * Cfg::W_BYTES, cv == 0 ? &tmW1 : &tmW2, &w_full[r.idx],
                                             tap * 128 + kc * 64, rank * 64);
                            r.advance(NW);
                        }
        }
    } else if (warp == 2) {
        if (rank == 0) { // ---------------- MMA issuer (leader CTA): whole warp, one elected lane per instruction
            constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
            Ring rw;
            uint32_t ph = 0;
            int it = 0;
            for (int lt = pair0; lt < n_loop; lt += pair_step, ph ^= 1, it++) {
                // Accumulator set `as` (256 TMEM columns) serves BOTH convolutions of this board: conv1 fills it, epilogue 1
                // drains it into slab H, conv2 refills it, epilogue 2 reads it - while the next board already runs both of
                // its convolutions in the other set.
                const int as = it & 1;
                const uint32_t sph = (uint32_t)((it >> 1) & 1);
                const uint32_t d0 = tmem_base + (uint32_t)(as * 2 * BN);
                for (int cv = 0; cv < 2; cv++) {
                    if (cv == 0) mbar_wait(&acc2_empty[as], sph ^ 1);  // epilogue 2 of board it-2 has drained this set
                    else mbar_wait(e1_done, ph);                       // epilogue 1: set drained, slab H written (both CTAs)
                    tc_fence_after();
                    for (int kc = 0; kc < 2; kc++) {
                        if (cv == 0) mbar_wait(&x_full[kc], ph);
                        const uint32_t slab_lo = umma_desc_lo(smem_u32((cv == 0 ? sX : sH) + kc * SLAB_BYTES) + (uint32_t)(HALO * 128));
                        int dy = -1, dx = -1;
                        for (int tap = 0; tap < 9; tap++) {
                            mbar_wait(&w_full[rw.idx], rw.phase);
                            tc_fence_after();
                            const uint32_t b_lo = umma_desc_lo(smem_u32(sW + rw.idx * Cfg::W_BYTES));
                            const uint32_t a_lo = slab_lo + (uint32_t)((dy * p.Wp + dx) * 8);
                            const uint32_t accf = (uint32_t)((kc | tap) != 0);
#pragma unroll
                            for (int sub = 0; sub < 2; sub++)
#pragma unroll
                                for (int k = 0; k < 4; k++)
                                    umma_bf16_elect<true>(d0 + (uint32_t)(sub * BN), a_lo + (uint32_t)(sub * 128 * 8 + k * 2),
                                                          b_lo + (uint32_t)(k * 2), idesc, k == 0 ? accf : 1u);
                            umma_commit_elect<true>(&w_empty[rw.idx]);
                            rw.advance(NW);
                            if (++dx == 2) { dx = -1; dy++; }
                        }
                        if (cv == 0) umma_commit_elect<true>(&x_empty[kc]);
                    }
                    if (cv == 0) umma_commit_elect<true>(acc1_full);
                    else { umma_commit_elect<true>(h_empty); umma_commit_elect<true>(&acc2_full[as]); }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) { // ---------------- epilogue 1: acc1 -> relu(BN2(conv1 + b1)) -> slab H (bf16, swizzled)
        const int q = warp & 3;
        uint32_t ph = 0;
        int it = 0;
        for (int lt = pair0; lt < n_loop; lt += pair_step, ph ^= 1, it++) {
            const int t = 2 * lt + rank;
            const int as = it & 1;
            mbar_wait(acc1_full, ph);
            mbar_wait(h_empty, ph ^ 1);  // conv2 of the previous board has finished reading slab H
            tc_fence_after();
            if (t < n_tiles) {
#pragma unroll 1
                for (int sub = 0; sub < 2; sub++) {
                    const int pos = sub * 128 + q * 32 + lane;
                    const int yy = pos / p.Wp, xx = pos - yy * p.Wp;
                    const bool live = yy != 0 && xx != p.Wp - 1;
                    const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * BN + sub * BN);
                    const int srow = HALO + pos;
#pragma unroll 1
                    for (int ck = 0; ck < 8; ck++) {
                        uint32_t r[16];
                        tmem_ld_32x16(t_acc + (uint32_t)(ck * 16), r);
                        tmem_ld_wait();
                        const float *pb = p.par1 + ck * 16, *sc = p.par1 + 128 + ck * 16, *sh = p.par1 + 256 + ck * 16;
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int c = 2 * i;
                            const float a0 = fmaxf(fmaf(sc[c], __uint_as_float(r[c]) + pb[c], sh[c]), 0.0f);
                            const float a1 = fmaxf(fmaf(sc[c + 1], __uint_as_float(r[c + 1]) + pb[c + 1], sh[c + 1]), 0.0f);
                            __nv_bfloat162 hh = __floats2bfloat162_rn(live ? a0 : 0.0f, live ? a1 : 0.0f);
                            w[i] = *reinterpret_cast<uint32_t *>(&hh);
                        }
                        // channels ck*16 .. +15 = K-block ck/4, 16-byte chunks 2*(ck%4) and +1, XOR-swizzled with the row phase
                        uint8_t *rowp = sH + (ck >> 2) * SLAB_BYTES + srow * 128;
                        const int j0 = (ck & 3) * 2;
                        *reinterpret_cast<uint4 *>(rowp + (((j0) ^ (srow & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4 *>(rowp + (((j0 + 1) ^ (srow & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(e1_done);
        }
    } else if (warp >= 8) { // ---------------- epilogue 2: SE + skip add + outputs (acc2)
        const int ew = warp - 8;
        const int q = warp & 3, sub = ew >> 2;
        const int et = threadIdx.x - 256; // 0..255
        float *s_part = s_se, *s_hp = s_se + 8 * BN, *s_gate = s_hp + 2 * BN;
        const float inv_cells = 1.0f / (float)p.n_cells;
        int it = 0;
        for (int lt = pair0; lt < n_loop; lt += pair_step, it++) {
            const int t = 2 * lt + rank;
            const int as = it & 1;
            const uint32_t sph = (uint32_t)((it >> 1) & 1);
            if (t >= n_tiles) { // dummy half of the last pair
                mbar_wait(&acc2_full[as], sph);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&acc2_empty[as]);
                continue;
            }
            const long long row = (long long)t * TILE_ROWS + sub * 128 + q * 32 + lane;
            const int pos = sub * 128 + q * 32 + lane;
            const int yy = pos / p.Wp, xx = pos - yy * p.Wp;
            const bool live = row < valid_rows && yy != 0 && xx != p.Wp - 1;
            const bool use_res = p.res && !(dbg & 8);
            float rnext[16];
            if (use_res) {
#pragma unroll 1
                for (int ck = 2; ck < 8; ck += 2) { // pull the rest of this row's residual into L2 meanwhile
                    const float *pp = p.res + f32_blk_index(row, ck * 16, BN);
#pragma unroll
                    for (int j = 0; j < 4; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + j * 256));
                }
                const size_t blk0 = f32_blk_index(row, 0, BN);
                ldg256(p.res + blk0, *reinterpret_cast<float(*)[8]>(&rnext[0]));
                ldg256(p.res + blk0 + 256, *reinterpret_cast<float(*)[8]>(&rnext[8]));
            }
            mbar_wait(&acc2_full[as], sph);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * BN + sub * BN);
            if (p.se && !(dbg & 64)) {
                // ---- pass 1: per-channel sums over the board's live cells
#pragma unroll 1
                for (int ck = 0; ck < 8; ck++) {
                    uint32_t r[16];
                    tmem_ld_32x16(t_acc + (uint32_t)(ck * 16), r);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) v[j] = live ? __uint_as_float(r[j]) : 0.0f;
                    int col = 0;
#pragma unroll
                    for (int m = 16, h = 8; m >= 2; m >>= 1, h >>= 1) {
                        const bool up = (lane & m) != 0;
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            if (i < h) {
                                const float send = up ? v[i] : v[i + h];
                                const float keep = up ? v[i + h] : v[i];
                                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                            }
                        col += up ? h : 0;
                    }
                    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
                    if ((lane & 1) == 0) s_part[(sub * 4 + q) * BN + ck * 16 + col] = v[0];
                }
                named_bar_sync(1, 256);
                {   // dense1 (C -> R) over 256 threads: output j, quarter `part` of the inputs
                    const int j = et & 63, part = et >> 6;
                    if (j < p.se_r) {
                        const float *w1 = p.se_w1 + (size_t)(part * 32) * p.se_r + j;
                        float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float m4[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int ii = part * 32 + i + u;
                                float sum = 0.0f;
#pragma unroll
                                for (int k = 0; k < 8; k++) sum += s_part[k * BN + ii];
                                m4[u] = sum * inv_cells;
                            }
                            h0 = fmaf(m4[0], __ldg(w1 + (size_t)(i) * p.se_r), h0);
                            h1 = fmaf(m4[1], __ldg(w1 + (size_t)(i + 1) * p.se_r), h1);
                            h2 = fmaf(m4[2], __ldg(w1 + (size_t)(i + 2) * p.se_r), h2);
                            h3 = fmaf(m4[3], __ldg(w1 + (size_t)(i + 3) * p.se_r), h3);
                        }
                        s_hp[part * 64 + j] = (h0 + h1) + (h2 + h3);
                    }
                }
                named_bar_sync(1, 256);
                {   // dense2 (R -> C): output channel cc, half `part` of the hidden units
                    const int cc = et & 127, part = et >> 7;
                    const int r2 = p.se_r >> 1;
                    const float *w2 = p.se_w2 + (size_t)(part * r2) * BN + cc;
                    float g0 = 0.0f, g1 = 0.0f;
#pragma unroll 16
                    for (int i = 0; i < r2; i += 2) {
                        const int ii = part * r2 + i;
                        const float ha = fmaxf(((s_hp[ii] + s_hp[64 + ii]) + (s_hp[128 + ii] + s_hp[192 + ii])) + __ldg(p.se_b1 + ii), 0.0f);
                        const float hb = fmaxf(((s_hp[ii + 1] + s_hp[65 + ii]) + (s_hp[129 + ii] + s_hp[193 + ii])) + __ldg(p.se_b1 + ii + 1), 0.0f);
                        g0 = fmaf(ha, __ldg(w2 + (size_t)(i) * BN), g0);
                        g1 = fmaf(hb, __ldg(w2 + (size_t)(i + 1) * BN), g1);
                    }
                    s_part[part * BN + cc] = g0 + g1;
                }
                named_bar_sync(1, 256);
                if (et < BN) s_gate[et] = 1.0f / (1.0f + expf(-((s_part[et] + s_part[BN + et]) + p.se_b2[et])));
                named_bar_sync(1, 256);
            } else if (p.se) {
                if (et < BN) s_gate[et] = 1.0f;
                named_bar_sync(1, 256);
            }
            // ---- output pass over 16-column chunks
#pragma unroll 1
            for (int ck = 0; ck < 8; ck++) {
                const int c0 = ck * 16;
                uint32_t r[16];
                tmem_ld_32x16(t_acc + (uint32_t)c0, r);
                float rcur[16];
                if (use_res) {
#pragma unroll
                    for (int j = 0; j < 16; j++) rcur[j] = rnext[j];
                    if (ck + 1 < 8) {
                        const size_t blk2 = f32_blk_index(row, c0 + 16, BN);
                        ldg256(p.res + blk2, *reinterpret_cast<float(*)[8]>(&rnext[0]));
                        ldg256(p.res + blk2 + 256, *reinterpret_cast<float(*)[8]>(&rnext[8]));
                    }
                }
                tmem_ld_wait();
                float v[16];
                const float *pb = p.par2 + c0;
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = __uint_as_float(r[j]) + pb[j];
                if (p.se) {
                    const float4 *g4 = reinterpret_cast<const float4 *>(s_gate + c0);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float4 g = g4[j];
                        v[4 * j] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
                    }
                }
                if (use_res) {
#pragma unroll
                    for (int j = 0; j < 16; j++) v[j] += rcur[j];
                }
                if (dbg & 4) continue;
                if (p.out_raw && !(dbg & 16)) {
                    const size_t blk = f32_blk_index(row, c0, BN);
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        float tt[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) tt[i] = live ? v[8 * j + i] : 0.0f;
                        stg256(p.out_raw + blk + j * 256, tt);
                    }
                }
#pragma unroll
                for (int o = 0; o < 2; o++) {
                    if (!(o == 0 ? p.out_a : p.out_b) || (dbg & 32)) continue;
                    const float *sc = p.par2 + (1 + 2 * o) * 128 + c0, *sh = p.par2 + (2 + 2 * o) * 128 + c0;
                    uint8_t *st = sStage + (ew * 2 + o) * 2048;
                    if ((ck & 1) == 0) { // the store that used this tile one chunk pair ago must have drained it
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int c = 8 * j + 2 * i;
                            const float a0 = fmaxf(fmaf(sc[c], v[c], sh[c]), 0.0f);
                            const float a1 = fmaxf(fmaf(sc[c + 1], v[c + 1], sh[c + 1]), 0.0f);
                            __nv_bfloat162 hh = __floats2bfloat162_rn(live ? a0 : 0.0f, live ? a1 : 0.0f);
                            w[i] = *reinterpret_cast<uint32_t *>(&hh);
                        }
                        const int piece = (ck & 1) * 2 + j; // 32-row x 32-channel SWIZZLE_64B tile: row = lane (64 B)
                        *reinterpret_cast<uint4 *>(st + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    if (ck & 1) {
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(o == 0 ? &tmOa : &tmOb, st, c0 - 16, (int)(row - lane));
                            tma_store_commit();
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acc2_empty[as]);
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

} // namespace gaz_block
