// gaz_conv.cuh -- board-tile implicit-GEMM convolution on tcgen05.
//
// One persistent CTA walks 256-row tiles of the padded-row activation tensor (Gomoku: 256 rows = exactly one
// board).  The kernel is built around the shared-memory port,
// which is what bounds an M=128 x N=128 cta_group::1 MMA (8 KB of operand reads per 64-cycle MMA = 128 B/clk):
//   * the activation SLAB of a tile (24 halo rows + 256 rows + 24 halo rows, 64 channels, SWIZZLE_128B) is
//     loaded ONCE per 64-channel K-block and all 9 filter taps read it through row-shifted UMMA descriptors
//     (start address + shift*128 B; measured on B200: the MMA unit applies the 128B swizzle to ABSOLUTE shared-
//     memory address bits exactly like TMA does, so a row-shifted start needs matrix-base-offset = 0 - setting it
//     to (addr >> 7) & 7 gives wrong results), so A is written to SMEM once, not 9x;
//   * every weight tile (tap, K-block) is used by both 128-row halves of the tile (8 MMAs per 16 KB tile);
//   * TMEM holds 2 stages x 2 accumulators (4 x BN columns) so the epilogue of tile i overlaps the MMAs of i+1.
// Warp roles (384 threads): 0 = slab TMA producer, 1 = weight TMA producer, 2 = MMA issuer (+TMEM alloc),
// 3 = idle, 4..11 = epilogue (TMEM lane quarter = warp & 3, 128-row half = (warp - 4) >> 2): two epilogue warps
// per scheduler so their TMEM / global / shuffle latencies overlap.
// Epilogue modes: plain (bias [+ fp32 residual] -> fp32 stream and/or relu(BN(.)) bf16 operands) and fused
// Squeeze-Excitation (Net/SE/SE_Block.py:15-23 + the block's skip add, Net/ResNet/ResNet_Block.py:27-41):
// pass 1 folds the accumulators into per-channel board means (halving butterfly over the 32 rows of a warp),
// the 128 epilogue threads run the two tiny dense layers + sigmoid, pass 2 re-reads TMEM and writes
// gate*(conv+bias) + residual.  fp32 row tensors use a 32x32-blocked layout (f32_blk_index) so that the
// row-per-lane accesses of the epilogue are 1 KB-contiguous 256-bit vectors instead of 32 separate lines.
#pragma once
#include "gaz_tc.cuh"
#include <cuda_bf16.h>

namespace gaz_conv {
using namespace gaz_tc;

// ---- blocked fp32 row tensor: element (row, c) of a [rows][C] tensor, rows % 32 == 0, C % 32 == 0 ----
// 32x32 blocks; inside a block 8-float pieces of a row are interleaved across the 32 rows so that lane r reading
// piece j of its row touches block + j*256 + r*8 floats: one warp instruction = 1 KB contiguous.
__host__ __device__ __forceinline__ size_t f32_blk_index(long long row, int c, int C) {
    const long long g = row >> 5;
    const int r = (int)(row & 31);
    const int cb = c >> 5, ci = c & 31;
    return ((size_t)(g * (C >> 5) + cb) << 10) + (size_t)((ci >> 3) << 8) + (size_t)(r << 3) + (size_t)(ci & 7);
}

__device__ __forceinline__ void ldg256(const float *p, float (&r)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(float *p, const float (&r)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
}
__device__ __forceinline__ void stg256u(void *p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// K-major SWIZZLE_128B descriptor whose start row is NOT aligned to the 8-row (1024 B) swizzle pattern.  The slab
// base is 1024 B aligned and was written by TMA, so the swizzle phase of every row is a function of its absolute
// address; base_offset_mode = 0 (default, parity-tested) leaves the matrix-base-offset field (bits 49..51) zero,
// mode 1 (debug) fills it with the phase of the start address.
__device__ __forceinline__ uint64_t umma_desc_sw128_shifted(uint32_t smem_addr, int base_offset_mode) {
    uint64_t d = umma_desc_sw128(smem_addr);
    if (base_offset_mode & 1) d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;
    return d;
}

struct BoardConvArgs {
    const int32_t *count;
    int max_count;
    int P_pad, Wp, H;         // rows per board (a divisor of 256 or (H+1)*Wp), row pitch W+1, board height: rows of a board beyond
                              // (H+1)*Wp are dead like the padding row / column
    int taps, kpt;            // filter taps (1 or 9; 3 in dx-merged mode), 64-channel K-blocks per tap
    int dxm;                  // dx-merged 3x3 convolution (BN = 96 / 48 instances only, see the kernel): taps = 3 row shifts dy * Wp
    // BN = 48: two small head convolutions on one input (8 + 8 output channels per dx), flat fp32 outputs [leaf][H*W][c]
    float *head_out[2];
    int head_c[2], W;
    int base_offset_mode;     // debug bits: 4 = skip the output stores (timing experiment)
    // per-channel parameters travel in the kernel-argument (constant) bank: bias | scale_a | shift_a | scale_b |
    // shift_b, 128 floats each.  The epilogue reads them with uniform LDC, which keeps them off the shared-memory
    // port that the MMAs saturate (a broadcast LDS per value cost 25 % of the port in v2.0, ncu r01).
    float par[5 * 128];
    const float *res;         // blocked fp32 residual stream (optional)
    float *out_raw;           // blocked fp32 output (optional)
    __nv_bfloat16 *out_a;     // relu(scale_a * v + shift_a) as bf16 rows (optional)
    __nv_bfloat16 *out_b;
    // dense-layer mode (heads): rows are leaves (all live), one launch per 128-wide slice of the outputs
    int dense;                // 1: every row < valid_rows is live; the padding mask is not applied
    int n_off;                // first output feature of this launch (row offset into the weight tensor / flat_out column)
    int n_slices;             // dense mode: 128-wide output slices computed by this launch (0 = 1); work items are (tile, slice)
                              // pairs, slice-minor, so the slices of a tile run on neighbouring CTA pairs at the same time and the
                              // tile's activations are read from HBM once; slice s uses par[s * 128 ..] as its biases
    float *flat_out;          // fp32 row-major [row][flat_ld] output (optional)
    int flat_ld, flat_n;      // leading dimension and number of valid output features
    // fused Squeeze-Excitation (tile == board, P_pad == 256): dense1 [C][R], dense2 [R][C]
    int se, se_r, n_cells;
    const float *se_w1, *se_b1, *se_w2, *se_b2;
};

constexpr int HALO = 24;                       // >= Wp + 1 for every game, multiple of 8
constexpr int TILE_ROWS = 256;
constexpr int SLAB_ROWS = TILE_ROWS + 2 * HALO; // 304 rows = 2 TMA boxes of 152 rows
constexpr int SLAB_BYTES = SLAB_ROWS * 128;     // 38912 = 38 * 1024
constexpr int SLAB_BOX_ROWS = SLAB_ROWS / 2;

template <int BN, bool PAIR> struct BoardCfg {
    static constexpr int NSLAB = PAIR ? 4 : 3;
    static constexpr int NB = BN <= 48 ? 16 : (BN <= 64 ? 8 : (BN == 96 ? 6 : 4)); // weight-tile ring: small tiles are consumed in
                                                                   // ~100 cycles each, the ring has to span the L2 latency
    static constexpr int NSTAGE = (PAIR && BN != 96 && BN != 48) ? 2 : 1; // 32-row x 32-channel bf16 staging tiles per epilogue warp (TMA store source)
    static constexpr int STAGE_BYTES = 8 * NSTAGE * 2048;
    static constexpr int B_ROWS = PAIR ? BN / 2 : BN; // weight rows (output channels) this CTA stages per tile
    static constexpr int B_BYTES = B_ROWS * 128;
    static constexpr int TMEM_COLS = BN == 96 ? 512 : (BN == 48 ? 256 : (4 * BN < 32 ? 32 : 4 * BN));   // a power of two
    static constexpr int SE_FLOATS = 8 * BN + BN + BN + BN; // partial sums [8 warps][BN], mean, hidden, gate
    static constexpr int SMEM = NSLAB * SLAB_BYTES + NB * B_BYTES + STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ +
                                SE_FLOATS * 4;
};

struct Ring {
    int idx;
    uint32_t phase;
    __device__ Ring() : idx(0), phase(0) {}
    __device__ void advance(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// PAIR = true: two CTAs of a cluster (one TPC) issue every MMA jointly (cta_group::2, M = 256 = one 128-row half of
// each CTA's own board tile); each CTA stages only half of every weight tile, which takes the shared-memory
// operand traffic per CTA from 8 KB to 6 KB per 64-cycle MMA - below the 128 B/clk port limit that caps the
// single-CTA form near 55 % of the tensor pipe (ncu r01).  Everything else (slab, TMEM, epilogue) stays per CTA.
template <int BN, bool PAIR>
__global__ void __launch_bounds__(384, 1)
conv_board_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOa, const __grid_constant__ CUtensorMap tmOb,
                  const __grid_constant__ BoardConvArgs p) {
    using Cfg = BoardCfg<BN, PAIR>;
    constexpr int NSLAB = Cfg::NSLAB, NB = Cfg::NB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;
    uint8_t *sB = base + NSLAB * SLAB_BYTES;
    uint8_t *sStage = sB + NB * Cfg::B_BYTES; // [8 epilogue warps][NSTAGE][2 KB], 1024 B aligned
    uint64_t *bars = (uint64_t *)(sStage + Cfg::STAGE_BYTES);
    uint64_t *a_full = bars, *a_empty = bars + NSLAB, *b_full = bars + 2 * NSLAB, *b_empty = bars + 2 * NSLAB + NB;
    uint64_t *tfull = bars + 2 * NSLAB + 2 * NB, *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);
    float *s_se = (float *)(bars + 64);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0; // 0 = leader (issues the MMAs, owns the full/tempty barriers)
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const long long valid_rows = (long long)cnt * p.P_pad;
    const int n_tiles = (int)((valid_rows + TILE_ROWS - 1) / TILE_ROWS);
    // loop index space: single CTA = tiles; pair = pair-tiles (tile 2*pt + rank; the last one may be a dummy)
    const int n_sl = p.n_slices > 1 ? p.n_slices : 1;
    const int n_loop = (PAIR ? (n_tiles + 1) / 2 : n_tiles) * n_sl;   // loop index = (pair-)tile * n_sl + slice
    const int dbg = p.base_offset_mode; // timing experiments: 8 no residual loads, 16 no fp32 stores, 32 no bf16 stores, 64 no SE passes

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLAB; s++) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], PAIR ? 16 : 8); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.out_a) tma_prefetch_desc(&tmOa);
        if (p.out_b) tma_prefetch_desc(&tmOb);
    }
    if (warp == 2) { if (PAIR) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS); else tmem_alloc(tmem_slot, Cfg::TMEM_COLS); }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) { // ---------------- activation-slab TMA producer
            Ring r;
            for (int lt = tile0; lt < n_loop; lt += tile_step) {
                const int tl = lt / n_sl;
                const int t = PAIR ? 2 * tl + rank : tl;
                const int row0 = t * TILE_ROWS - HALO;
                for (int kc = 0; kc < p.kpt; kc++) {
                    mbar_wait(&a_empty[r.idx], r.phase ^ 1);
                    uint8_t *dst = sA + r.idx * SLAB_BYTES;
                    if (PAIR) {
                        if (rank == 0) mbar_expect_tx(&a_full[r.idx], 2 * SLAB_BYTES);
                        tma_load_2d_pair(dst, &tmA, &a_full[r.idx], kc * 64, row0);
                        tma_load_2d_pair(dst + SLAB_BYTES / 2, &tmA, &a_full[r.idx], kc * 64, row0 + SLAB_BOX_ROWS);
                    } else {
                        mbar_expect_tx(&a_full[r.idx], SLAB_BYTES);
                        tma_load_2d(dst, &tmA, &a_full[r.idx], kc * 64, row0);
                        tma_load_2d(dst + SLAB_BYTES / 2, &tmA, &a_full[r.idx], kc * 64, row0 + SLAB_BOX_ROWS);
                    }
                    r.advance(NSLAB);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) { // ---------------- weight-tile TMA producer: order (K-block, tap)
            Ring r;
            const int cin = p.kpt * 64;
            for (int lt = tile0; lt < n_loop; lt += tile_step) {
                const int n_off = p.n_off + (lt % n_sl) * 128;
                for (int kc = 0; kc < p.kpt; kc++)
                    for (int tap = 0; tap < p.taps; tap++) {
                        mbar_wait(&b_empty[r.idx], r.phase ^ 1);
                        if (PAIR) {
                            if (rank == 0) mbar_expect_tx(&b_full[r.idx], 2 * Cfg::B_BYTES);
                            tma_load_2d_pair(sB + r.idx * Cfg::B_BYTES, &tmB, &b_full[r.idx], tap * cin + kc * 64, n_off + rank * Cfg::B_ROWS);
                        } else {
                            mbar_expect_tx(&b_full[r.idx], Cfg::B_BYTES);
                            tma_load_2d(sB + r.idx * Cfg::B_BYTES, &tmB, &b_full[r.idx], tap * cin + kc * 64, n_off);
                        }
                        r.advance(NB);
                    }
            }
        }
    } else if (warp == 2) {
        if (rank == 0) { // ---------------- MMA issuer: whole warp, uniform control flow, one elected lane per instruction
            constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN);
            Ring ra, rb;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int lt = tile0; lt < n_loop; lt += tile_step) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * 2 * BN);
                for (int kc = 0; kc < p.kpt; kc++) {
                    mbar_wait(&a_full[ra.idx], ra.phase);
                    const uint32_t slab_lo = umma_desc_lo(smem_u32(sA + ra.idx * SLAB_BYTES) + (uint32_t)(HALO * 128));
                    int dy = -1, dx = -1; // tap = (dy + 1) * 3 + (dx + 1)
                    for (int tap = 0; tap < p.taps; tap++) {
                        mbar_wait(&b_full[rb.idx], rb.phase);
                        tc_fence_after();
                        const int shift = (BN == 96 || BN == 48) ? (tap - 1) * p.Wp : (p.taps == 9 ? dy * p.Wp + dx : 0);
                        const uint32_t b_lo = umma_desc_lo(smem_u32(sB + rb.idx * Cfg::B_BYTES));
                        const uint32_t a_lo = slab_lo + (uint32_t)(shift * 8); // 128 B per row = 8 descriptor units
                        const uint32_t accf = (uint32_t)((kc | tap) != 0); // the first MMA into each accumulator overwrites
#pragma unroll
                        for (int sub = 0; sub < 2; sub++) {
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                umma_bf16_elect<PAIR>(d0 + (uint32_t)(sub * BN), a_lo + (uint32_t)(sub * 128 * 8 + k * 2),
                                                      b_lo + (uint32_t)(k * 2), idesc, k == 0 ? accf : 1u);
                            }
                        }
                        umma_commit_elect<PAIR>(&b_empty[rb.idx]);
                        rb.advance(NB);
                        if (++dx == 2) { dx = -1; dy++; }
                    }
                    umma_commit_elect<PAIR>(&a_empty[ra.idx]);
                    ra.advance(NSLAB);
                }
                umma_commit_elect<PAIR>(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) { // ---------------- epilogue warps: TMEM lane quarter q, 128-row half `sub`
        const int q = warp & 3, sub = (warp - 4) >> 2;
        const int et = threadIdx.x - 128; // 0..255
        float *s_part = s_se, *s_mean = s_se + 8 * BN, *s_hid = s_mean + BN, *s_gate = s_hid + BN;
        constexpr int NCH = BN / 32;
        uint32_t stage_i = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int lt = tile0; lt < n_loop; lt += tile_step) {
            const int tl = lt / n_sl, sl = lt - tl * n_sl;
            const int n_off = p.n_off + sl * 128;
            const int t = PAIR ? 2 * tl + rank : tl;
            if (t >= n_tiles) { // dummy half of the last pair-tile: keep the barrier phases in step, touch nothing
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            const long long row = (long long)t * TILE_ROWS + sub * 128 + q * 32 + lane;
            const int pos = (int)(row % p.P_pad);
            const int yy = pos / p.Wp, xx = pos - yy * p.Wp;
            const bool live = row < valid_rows && (p.dense || (yy != 0 && yy <= p.H && xx != p.Wp - 1));
            if constexpr (BN == 48) {
                // ---- dx-merged pair of small head convolutions (Connect4: policy and value head, 3x3 C128 -> C8 each, on the
                // bf16 copy of the trunk output): as BN = 96 below with 16 columns per dx (head 0: columns 0..7, head 1: 8..15);
                // the sums go out as flat fp32 [leaf][cell][c] rows, the layout the dense stack behind a head reads.
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * 2 + sub) * BN);
                uint32_t em[16], e0[16], ep[16];
                tmem_ld_32x16(t_acc, em);
                tmem_ld_32x16(t_acc + 16u, e0);
                tmem_ld_32x16(t_acc + 32u, ep);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    float a = __shfl_up_sync(0xffffffffu, __uint_as_float(em[c]), 1);
                    float b = __shfl_down_sync(0xffffffffu, __uint_as_float(ep[c]), 1);
                    if (lane == 0) a = 0.0f;
                    if (lane == 31) b = 0.0f;
                    v[c] = ((a + __uint_as_float(e0[c])) + b) + p.par[c];
                }
                if (live) {
                    const int board = (int)(row / p.P_pad), cell = (yy - 1) * p.W + xx;
#pragma unroll
                    for (int hd = 0; hd < 2; hd++) {
                        if (!p.head_out[hd]) continue;
                        float *o = p.head_out[hd] + ((size_t)board * (size_t)(p.H * p.W) + (size_t)cell) * (size_t)p.head_c[hd];
                        if (p.head_c[hd] == 8) stg256(o, *reinterpret_cast<const float(*)[8]>(&v[8 * hd]));
                        else
#pragma unroll
                            for (int c = 0; c < 8; c++)
                                if (c < p.head_c[hd]) o[c] = v[8 * hd + c];
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            if constexpr (BN == 96) {
                // ---- dx-merged 3x3 convolution with 32 outputs (the C128 -> C32 head convolutions).  The weight tile of row
                // shift dy holds the filters of its three taps side by side (N = 96: column dx' * 32 + c, dx' = dx + 1), so
                // an MMA computes E_dx[r][c] = sum_dy sum_k A[r + dy * Wp][k] * W[dy][dx][k][c] for all three dx from ONE
                // read of A (a N = 32 MMA reads the same 4 KB of A for a third of the work and is bound by the shared-
                // memory port), and the output is out[r] = E_-1[r - 1] + E_0[r] + E_+1[r + 1]: the two neighbour terms come
                // from the adjacent TMEM lanes by warp shuffle.  Wp divides 32 and a tile starts with a board, so the
                // lanes 0 / 31 of a warp are cells of column 0 / the padding column: the term a lane 0 would need from the
                // previous warp belongs to a padding cell (its A rows are zero: E = 0 exactly), lane 31's output is dead.
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * 2 + sub) * BN);
                uint8_t *st = sStage + (warp - 4) * Cfg::NSTAGE * 2048;
#pragma unroll 1
                for (int hf = 0; hf < 2; hf++) {
                    uint32_t em[16], e0[16], ep[16];
                    tmem_ld_32x16(t_acc + (uint32_t)(hf * 16), em);
                    tmem_ld_32x16(t_acc + (uint32_t)(32 + hf * 16), e0);
                    tmem_ld_32x16(t_acc + (uint32_t)(64 + hf * 16), ep);
                    tmem_ld_wait();
                    if (hf == 1) { // the accumulator is in registers: hand it back to the MMA issuer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
                    } else {       // the previous tile's store must have drained the staging tile
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                    }
                    const float *pb = p.par + hf * 16, *sc = p.par + 128 + hf * 16, *sh = p.par + 256 + hf * 16;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            float o2[2];
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                const int c = 8 * j + 2 * i + u;
                                float a = __shfl_up_sync(0xffffffffu, __uint_as_float(em[c]), 1);
                                float b = __shfl_down_sync(0xffffffffu, __uint_as_float(ep[c]), 1);
                                if (lane == 0) a = 0.0f;
                                if (lane == 31) b = 0.0f;
                                const float v = ((a + __uint_as_float(e0[c])) + b) + pb[c];
                                o2[u] = live ? fmaxf(fmaf(sc[c], v, sh[c]), 0.0f) : 0.0f;
                            }
                            __nv_bfloat162 hh = __floats2bfloat162_rn(o2[0], o2[1]);
                            w[i] = *reinterpret_cast<uint32_t *>(&hh);
                        }
                        const int piece = hf * 2 + j;
                        *reinterpret_cast<uint4 *>(st + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmOa, st, 0, (int)(row - lane));
                    tma_store_commit();
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            float rnext[32]; // residual of the first output chunk: issued before the accumulator is even ready
            const bool use_res = p.res && !(dbg & 8);
            if (use_res) {
#pragma unroll 1
                for (int ch = 1; ch < NCH; ch++) { // pull the rest of this row's residual into L2 meanwhile
                    const float *pp = p.res + f32_blk_index(row, ch * 32, BN);
#pragma unroll
                    for (int j = 0; j < 4; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + j * 256));
                }
                const size_t blk0 = f32_blk_index(row, 0, BN);
#pragma unroll
                for (int j = 0; j < 4; j++) ldg256(p.res + blk0 + j * 256, *reinterpret_cast<float(*)[8]>(&rnext[8 * j]));
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * 2 + sub) * BN);
            if (p.se && (dbg & 64)) { if (et < BN) s_gate[et] = 1.0f; named_bar_sync(1, 256); }
            if (p.se && !(dbg & 64)) {
                // ---- pass 1: per-channel sums over the board's live cells (tile == board)
#pragma unroll 1
                for (int ch = 0; ch < ((dbg & 256) ? 0 : NCH); ch++) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_acc + (uint32_t)(ch * 32), r);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = live ? __uint_as_float(r[j]) : 0.0f;
                    // halving butterfly: 31 shuffles leave lane l with the column sum of column f(l)
                    int col = 0;
#pragma unroll
                    for (int m = 16, h = 16; m >= 1; m >>= 1, h >>= 1) {
                        const bool up = (lane & m) != 0;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            if (i < h) {
                                const float send = up ? v[i] : v[i + h];
                                const float keep = up ? v[i + h] : v[i];
                                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                            }
                        }
                        col += up ? h : 0;
                    }
                    s_part[(sub * 4 + q) * BN + ch * 32 + col] = v[0];
                }
                named_bar_sync(1, 256);
                if (!(dbg & 128)) {
                    // dense1 (C -> R) spread over all 256 epilogue threads: output j, quarter `part` of the inputs, all
                    // 32 weight loads in flight at once (the weights live in L2: with 64 / 128 threads and 128 / 64
                    // dependent steps the two tiny layers cost 0.38 ms per launch, ncu + ablation r01).  The conv bias
                    // is folded into the dense bias on the host (se_b1 = b1 + W1^T bias).
                    const int j = et & 63, part = et >> 6;
                    if (j < p.se_r) {
                        const float *w1 = p.se_w1 + (size_t)(part * (BN / 4)) * p.se_r + j;
                        const float inv_cells = 1.0f / (float)p.n_cells;
                        float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
                        for (int i = 0; i < BN / 4; i += 4) {
                            float m4[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int ii = part * (BN / 4) + i + u;
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
                        s_mean[part * 64 + j] = (h0 + h1) + (h2 + h3);   // hidden partials [4][64] (s_mean .. s_hid)
                    }
                }
                named_bar_sync(1, 256);
                if (!(dbg & 128)) {
                    // dense2 (R -> C): output channel cc, half `part` of the hidden units (relu(dense1) rebuilt on the fly)
                    const int cc = et & (BN - 1), part = et / BN;
                    const int r2 = p.se_r >> 1;
                    if (part < 2) {
                        const float *w2 = p.se_w2 + (size_t)(part * r2) * BN + cc;
                        float g0 = 0.0f, g1 = 0.0f;
#pragma unroll 16
                        for (int i = 0; i < r2; i += 2) {
                            const int ii = part * r2 + i;
                            const float ha = fmaxf(((s_mean[ii] + s_mean[64 + ii]) + (s_mean[128 + ii] + s_mean[192 + ii])) + __ldg(p.se_b1 + ii), 0.0f);
                            const float hb = fmaxf(((s_mean[ii + 1] + s_mean[65 + ii]) + (s_mean[129 + ii] + s_mean[193 + ii])) + __ldg(p.se_b1 + ii + 1), 0.0f);
                            g0 = fmaf(ha, __ldg(w2 + (size_t)(i) * BN), g0);
                            g1 = fmaf(hb, __ldg(w2 + (size_t)(i + 1) * BN), g1);
                        }
                        s_part[part * BN + cc] = g0 + g1;                  // gate partials reuse the sums' slots
                    }
                }
                named_bar_sync(1, 256);
                if (et < BN) s_gate[et] = (dbg & 128) ? 0.5f : 1.0f / (1.0f + expf(-((s_part[et] + s_part[BN + et]) + p.se_b2[et])));
                named_bar_sync(1, 256);
            }
            // ---- output pass over the 32-column chunks of this warp's 32 rows; the residual of chunk i+1 is
            // loaded while chunk i is processed (register double buffer)
#pragma unroll 1
            for (int ch = 0; ch < NCH; ch++) {
                uint32_t r[32];
                tmem_ld_32x32(t_acc + (uint32_t)(ch * 32), r);
                float rcur[32];
                if (use_res) {
#pragma unroll
                    for (int j = 0; j < 32; j++) rcur[j] = rnext[j];
                    if (ch + 1 < NCH) {
                        const size_t blk2 = f32_blk_index(row, (ch + 1) * 32, BN);
#pragma unroll
                        for (int j = 0; j < 4; j++) ldg256(p.res + blk2 + j * 256, *reinterpret_cast<float(*)[8]>(&rnext[8 * j]));
                    }
                }
                tmem_ld_wait();
                float v[32];
                const float *pb = p.par + sl * 128 + ch * 32; // constant bank, uniform index: no shared-memory traffic
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]) + pb[j];
                if (p.se) {
                    const float4 *g4 = reinterpret_cast<const float4 *>(s_gate + ch * 32);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 g = g4[j];
                        v[4 * j] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
                    }
                }
                if (use_res) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] += rcur[j];
                }
                const size_t blk = f32_blk_index(row, ch * 32, BN); // this lane's first piece of the 32x32 block
                if (dbg & 4) continue; // timing experiment only: no output stores
                if (p.flat_out && live) { // dense layers: plain row-major fp32 [leaf][features]
                    float *fo = p.flat_out + (size_t)row * p.flat_ld + n_off + ch * 32;
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (n_off + ch * 32 + j < p.flat_n) fo[j] = v[j];
                }
                if (p.out_raw && !(dbg & 16)) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float tt[8];
#pragma unroll
                        for (int i = 0; i < 8; i++) tt[i] = live ? v[8 * j + i] : 0.0f;
                        stg256(p.out_raw + blk + j * 256, tt);
                    }
                }
                // bf16 operands of the next layer: the 32 x 32 tile of this warp goes through a 2 KB SWIZZLE_64B
                // staging tile and leaves with ONE TMA store - a row-per-lane STG touches 32 lines per instruction
                // and made the epilogue LSU-bound (ncu r01)
#pragma unroll
                for (int o = 0; o < 2; o++) {
                    if (!(o == 0 ? p.out_a : p.out_b) || (dbg & 32)) continue;
                    const float *sc = p.par + (1 + 2 * o) * 128 + ch * 32, *sh = p.par + (2 + 2 * o) * 128 + ch * 32;
                    uint8_t *st = sStage + ((warp - 4) * Cfg::NSTAGE + (stage_i % Cfg::NSTAGE)) * 2048;
                    stage_i++;
                    // at most NSTAGE - 1 earlier stores of this warp may still be reading their staging tiles
                    if (lane == 0) { if (Cfg::NSTAGE == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int c = 8 * j + 2 * i;
                            const float a0 = fmaxf(fmaf(sc[c], v[c], sh[c]), 0.0f);
                            const float a1 = fmaxf(fmaf(sc[c + 1], v[c + 1], sh[c + 1]), 0.0f);
                            __nv_bfloat162 hh = __floats2bfloat162_rn(live ? a0 : 0.0f, live ? a1 : 0.0f);
                            w[i] = *reinterpret_cast<uint32_t *>(&hh);
                        }
                        // row = lane (64 B), 16-byte chunk j XOR-swizzled with address bits 7..8 (= (lane >> 1) & 3)
                        *reinterpret_cast<uint4 *>(st + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(o == 0 ? &tmOa : &tmOb, st, ch * 32, (int)(row - lane));
                        tma_store_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    if (warp >= 4 && lane == 0) tma_store_wait_all(); // bulk stores must complete before the CTA's shared memory goes away
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) { if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

} // namespace gaz_conv
