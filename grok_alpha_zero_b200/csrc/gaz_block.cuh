// gaz_block.cuh -- one pre-activation residual block (Net/ResNet/ResNet_Block.py:27-41 + Net/SE/SE_Block.py:15-23) as ONE
// kernel for tile == board geometries (Gomoku: 256 padded rows per board), C_in = 128 or 256, C_out = 128:
//
//   a_x (bf16, HBM) --TMA--> slab --conv1 MMAs--> TMEM --epilogue 1: relu(BN2(. + b1))--> the SAME slab (bf16, in place)
//        --conv2 MMAs--> TMEM --epilogue 2: SE gate * (. + b2) + residual--> x_out (fp32) and relu(BN(x_out)) (bf16)
//
// The bf16 intermediate between the two convolutions never leaves the SM (2.1 GB per block and launch at 16384 leaves),
// and conv1 - a tensor-bound kernel on its own - runs in the shadow of the memory-bound fused-SE epilogue of the
// previous board.  Every board depends only on itself: the rows a live output reads outside its own board are padding
// rows/columns, which are zero in the input slab's halo as well, so conv1's output can overwrite the input slab in place
// (rows 0..255 only; the halo keeps the neighbours' padding zeros) - the shared memory this saves is the 12-deep weight
// ring that keeps the tensor pipe fed across L2 latency.  C_in = 256 (first block of the trunk): the four input K-blocks
// stream through the two slab halves (K-block kc -> half kc & 1).
// CTA pairs (cta_group::2) as in gaz_conv.cuh: each CTA owns its board, slab, TMEM and epilogues; the leader issues
// every MMA for both (M = 256) and each CTA stages half of every weight tile.
// Warp roles (512 threads): 0 slab TMA producer, 1 weight TMA producer (W1 tiles then W2 tiles per board), 2 MMA
// issuer, 3 idle, 4..7 epilogue 1 (one per TMEM lane quarter), 8..15 epilogue 2 (quarter x 128-row half).
// TMEM: two accumulator sets of 256 columns (2 x 128-row halves x 128 channels); boards alternate between the sets and
// BOTH convolutions of a board use the board's set (conv1 -> epilogue 1 drains it into the slab -> conv2 -> epilogue 2),
// so conv1, epilogue 1 and conv2 of board i+1 all run while epilogue 2 of board i is still reading the other set.
// Both epilogues are software-pipelined (the TMEM load / residual load of the next piece is in flight while the current
// one is converted) and keep their per-channel parameters out of registers (shared-memory broadcast loads, indexed
// constant bank).  Cycle accounting per role: build with -DGAZ_BLOCK_CLK, run with debug bit 1024.
//
// TRUNK LAUNCHES.  A launch runs up to MAX_LAYERS consecutive residual blocks: every CTA walks its tiles once per layer
// (layer-major), so tile t of layer l + 1 is processed by the CTA that produced it in layer l, n_my items earlier.  Boards
// depend only on themselves, so the only cross-layer ordering is inside a CTA: the slab producer waits until the epilogue-2
// warps have seen their TMA stores of (l, t) complete (per-warp progress counters in shared memory) before it loads (l + 1, t);
// the fp32 residual is written and re-read by the same thread.  The pipeline (weight ring, slab prefetch, accumulator
// ping-pong) never drains between layers: the ~25 us of fill / drain and launch gap that every separate block launch pays
// (a fifth of a 4096-leaf Connect4 block) is paid once per launch.
#pragma once
#include "gaz_conv.cuh"

#ifdef GAZ_BLOCK_CLK   // per-role stall accounting (clock64 + printf from CTA 0), build with -DGAZ_BLOCK_CLK, run with dbg bit 1024
#define TK_BEGIN() do { if (dbg & 1024) tq = clock64(); } while (0)
#define TK_END(acc) do { if (dbg & 1024) acc += clock64() - tq; } while (0)
#define TK_PRINT(...) printf(__VA_ARGS__)
#define TK_START(v) v = clock64()
#define TK_STOP(v) v = clock64() - v
#else
#define TK_START(v) do { } while (0)
#define TK_STOP(v) do { } while (0)
#define TK_BEGIN() do { } while (0)
#define TK_END(acc) do { } while (0)
#define TK_PRINT(...) do { } while (0)
#endif

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

constexpr int MAX_LAYERS = 6;   // kernel parameters: 6 x 4.8 KB < 32 KB

struct TrunkLayer {           // one residual block
    CUtensorMap tmA, tmW1, tmW2, tmOa, tmOb;   // input operand | conv1 / conv2 filters | bf16 outputs
    float par1[3 * 128];      // conv1 bias | BN2 scale | BN2 shift          (constant bank, uniform loads)
    float par2[5 * 128];      // conv2 bias | scale_a | shift_a | scale_b | shift_b
    const float *res;         // blocked fp32 residual stream in
    float *out_raw;           // blocked fp32 residual stream out
    __nv_bfloat16 *out_a, *out_b;
    int nkc1;                 // 64-channel K-blocks of conv1's input (2: C_in = 128, 4: C_in = 256)
    int dep;                  // the input operand / residual of this layer are outputs of the previous layer of this launch
    int plain_a;              // out_a = bf16(scale_a * v + shift_a) WITHOUT the ReLU (the bf16 copy of the trunk output that a
                              // tensor-core head convolution reads, Connect4/Build_Model.py:27,48)
    int se, se_r;
    const float *se_w1, *se_b1, *se_w2, *se_b2; // se_b1 has the conv2 bias folded in (b1 + W1^T bias2)
};

struct TrunkArgs {
    const int32_t *count;
    int max_count;
    int Wp, H, P_pad, n_cells, dbg;   // P_pad divides 256: a tile is 256 / P_pad whole boards (Gomoku 1, Connect4 4, TicTacToe 16)
    int n_layers;
    TrunkLayer L[MAX_LAYERS];
};

struct Cfg {
    static constexpr int NW = 12;                   // weight-tile ring (half tiles: 64 output channels x 64 k)
    static constexpr int W_BYTES = 64 * 128;
    static constexpr int STAGE_BYTES = 8 * 2 * 2048; // epilogue-2 warps: one 32 x 32-channel bf16 tile per output
    static constexpr int SE_FLOATS = 8 * 128 + 128 + 256 + 64 + 128 + 128;
    static constexpr int SMEM = 2 * SLAB_BYTES + NW * W_BYTES + STAGE_BYTES + 1024 + 512 + SE_FLOATS * 4 + 2 * 128 * 4 + 64;
};

// per-channel sums of 32 accumulator columns over the 32 rows of a warp: halving butterfly (31 shuffles), lane l ends
// with the sum of column `col(l)` and stores it to dst[col]
__device__ __forceinline__ void colsum32(const uint32_t (&r)[32], uint32_t mask, int lane, float *dst) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j] & mask);
    int col = 0;
#pragma unroll
    for (int m = 16, h = 16; m >= 1; m >>= 1, h >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i < h) {
                const float send = up ? v[i] : v[i + h];
                const float keep = up ? v[i + h] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
        col += up ? h : 0;
    }
    dst[col] = v[0];
}

__device__ __forceinline__ uint32_t pack_plain_bf16x2(float lo, float hi) { // {hi, lo} -> bf16x2
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) { // {hi, lo} -> max(., 0) -> bf16x2
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// epilogue 2, one 16-column chunk of one row: gate * (conv2 + bias) + residual -> fp32 stream, relu(BN(.)) -> bf16
// operand(s) of the next layer through the warp's 32 x 32-channel SWIZZLE_64B staging tile (one TMA store per 2 chunks)
template <int ODD>
__device__ __forceinline__ void e2_chunk(const uint32_t (&acc)[16], const float (&res)[16], bool use_res, int ck, const TrunkLayer &p,
                                         float *outp, uint32_t gate_addr, uint32_t bg_addr, uint32_t stage_addr, uint32_t mask,
                                         int lane, int row0, int dbg) {
    const int c0 = ck * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float4 g = lds128f(gate_addr + (uint32_t)(c0 * 4 + j * 16)), b = lds128f(bg_addr + (uint32_t)(c0 * 4 + j * 16));
        v[4 * j] = fmaf(__uint_as_float(acc[4 * j]), g.x, b.x);
        v[4 * j + 1] = fmaf(__uint_as_float(acc[4 * j + 1]), g.y, b.y);
        v[4 * j + 2] = fmaf(__uint_as_float(acc[4 * j + 2]), g.z, b.z);
        v[4 * j + 3] = fmaf(__uint_as_float(acc[4 * j + 3]), g.w, b.w);
    }
    if (use_res) {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] += res[j];
    }
    if (dbg & 4) return;
    if (p.out_raw && !(dbg & 16)) { // padding rows / columns carry don't-care values in the fp32 stream (never read by a live cell)
        stg256(outp + (2 * ck) * 256, *reinterpret_cast<const float(*)[8]>(&v[0]));
        stg256(outp + (2 * ck + 1) * 256, *reinterpret_cast<const float(*)[8]>(&v[8]));
    }
#pragma unroll
    for (int o = 0; o < 2; o++) {
        if (!(o == 0 ? p.out_a : p.out_b) || (dbg & 32)) continue;
        const float *sc = p.par2 + (1 + 2 * o) * 128 + c0, *sh = p.par2 + (2 + 2 * o) * 128 + c0;
        const uint32_t st = stage_addr + (uint32_t)(o * 2048);
        if (!ODD) { // the store that used this tile one chunk pair ago must have drained it
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = 8 * j + 2 * i;
                const float f0 = fmaf(sc[c], v[c], sh[c]), f1 = fmaf(sc[c + 1], v[c + 1], sh[c + 1]);
                w[i] = ((o == 0 && p.plain_a) ? pack_plain_bf16x2(f0, f1) : pack_relu_bf16x2(f0, f1)) & mask;
            }
            const int piece = ODD * 2 + j; // 32-row x 32-channel SWIZZLE_64B tile: row = lane (64 B)
            sts128(st + (uint32_t)(lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4)), w[0], w[1], w[2], w[3]);
        }
        if (ODD) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d_addr(o == 0 ? &p.tmOa : &p.tmOb, st, c0 - 16, row0);
                tma_store_commit();
            }
        }
    }
}

// epilogue 1, one piece: 32 accumulator columns (channels c32*32 .. +31) of one row -> relu(BN2(. + b1)) -> bf16 -> the
// row's four 16-byte chunks in K-block c32/2 of slab H, XOR-swizzled with the row phase like a SWIZZLE_128B TMA write.
// The per-channel scale and (bias-folded) shift come from shared memory through volatile broadcast loads: as kernel
// arguments the compiler hoists them across pieces into registers and spills them.
__device__ __forceinline__ void e1_piece(const uint32_t (&r)[32], uint32_t par_addr, uint32_t sH_addr, int c32, int srow, bool live) {
    const uint32_t rowp = sH_addr + (uint32_t)((c32 >> 1) * SLAB_BYTES + srow * 128);
    const int sw = srow & 7, j0 = (c32 & 1) * 4;
    const uint32_t mask = live ? 0xffffffffu : 0u;   // padding rows / columns of the board stay zero
    const uint32_t sc_addr = par_addr + (uint32_t)(c32 * 128), sh_addr = sc_addr + 512;
    // the parameters of chunk ch + 1 are loaded before chunk ch is converted (the volatile loads / stores keep their program
    // order, so without this the 30-cycle shared-memory latency is exposed once per chunk)
    float4 s0 = lds128f(sc_addr), s1 = lds128f(sc_addr + 16), h0 = lds128f(sh_addr), h1 = lds128f(sh_addr + 16);
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
        float4 ns0 = s0, ns1 = s1, nh0 = h0, nh1 = h1;
        if (ch < 3) {
            ns0 = lds128f(sc_addr + (ch + 1) * 32); ns1 = lds128f(sc_addr + (ch + 1) * 32 + 16);
            nh0 = lds128f(sh_addr + (ch + 1) * 32); nh1 = lds128f(sh_addr + (ch + 1) * 32 + 16);
        }
        const int j = ch * 8;
        uint32_t w[4];
        w[0] = pack_relu_bf16x2(fmaf(s0.x, __uint_as_float(r[j]), h0.x), fmaf(s0.y, __uint_as_float(r[j + 1]), h0.y)) & mask;
        w[1] = pack_relu_bf16x2(fmaf(s0.z, __uint_as_float(r[j + 2]), h0.z), fmaf(s0.w, __uint_as_float(r[j + 3]), h0.w)) & mask;
        w[2] = pack_relu_bf16x2(fmaf(s1.x, __uint_as_float(r[j + 4]), h1.x), fmaf(s1.y, __uint_as_float(r[j + 5]), h1.y)) & mask;
        w[3] = pack_relu_bf16x2(fmaf(s1.z, __uint_as_float(r[j + 6]), h1.z), fmaf(s1.w, __uint_as_float(r[j + 7]), h1.w)) & mask;
        sts128(rowp + (uint32_t)(((j0 + ch) ^ sw) << 4), w[0], w[1], w[2], w[3]);
        s0 = ns0; s1 = ns1; h0 = nh0; h1 = nh1;
    }
}

// epilogue 1 for one 128-row half of a board: this warp's 32 rows x 128 channels, four 32-column pieces, the TMEM load of
// the next piece in flight while the current one is converted
__device__ __forceinline__ void e1_drain(uint32_t t_sub, uint32_t par_addr, uint32_t sH_addr, int srow, bool live) {
    uint32_t ra[32], rb[32];
    tmem_ld_32x32(t_sub, ra);
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
        tmem_ld_wait_dep(ra);
        tmem_ld_32x32(t_sub + (uint32_t)(c * 64 + 32), rb);
        e1_piece(ra, par_addr, sH_addr, 2 * c, srow, live);
        tmem_ld_wait_dep(rb);
        if (c == 0) tmem_ld_32x32(t_sub + 64u, ra);
        e1_piece(rb, par_addr, sH_addr, 2 * c + 1, srow, live);
    }
}
// epilogue 1 of one 128-row half for a whole warp: wait for that half of conv1, drain, publish the slab rows to the async
// proxy, arrive
__device__ __forceinline__ void e1_half(uint64_t *acc1_full, uint64_t *e1_done, uint32_t ph, bool work, uint32_t t_sub,
                                        uint32_t par_addr, uint32_t sH_addr, int srow, bool live, int lane) {
    mbar_wait(acc1_full, ph);
    tc_fence_after();
    if (work) e1_drain(t_sub, par_addr, sH_addr, srow, live);
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_leader(e1_done);
}

// MMAs of one (K-block, tap) weight tile for the 128-row halves SUBS (bit 0: rows 0..127, bit 1: rows 128..255 of each
// CTA's board), then the tile goes back to the weight producer
template <int SUBS>
__device__ __forceinline__ void issue_tile(uint32_t d0, uint32_t slab_lo, uint32_t b_lo, int tap, int Wp, uint32_t accf, uint32_t idesc) {
    const int dy = tap / 3 - 1, dx = tap - (dy + 1) * 3 - 1;
    const uint32_t a_lo = slab_lo + (uint32_t)((dy * Wp + dx) * 8);
#pragma unroll
    for (int sub = 0; sub < 2; sub++) {
        if (!((SUBS >> sub) & 1)) continue;
#pragma unroll
        for (int k = 0; k < 4; k++)
            umma_bf16_elect<true>(d0 + (uint32_t)(sub * 128), a_lo + (uint32_t)(sub * 128 * 8 + k * 2), b_lo + (uint32_t)(k * 2), idesc,
                                  k == 0 ? accf : 1u);
    }
}

__global__ void __launch_bounds__(512, 1)
res_trunk_kernel(const __grid_constant__ TrunkArgs p) {
    constexpr int BN = 128, NW = Cfg::NW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sX = base;                       // 2 slabs: K-blocks 0/1 of the block input, then (in place) of the conv1 output
    uint8_t *sH = sX;
    uint8_t *sW = base + 2 * SLAB_BYTES;
    uint8_t *sStage = sW + NW * Cfg::W_BYTES;
    uint64_t *bars = (uint64_t *)(sStage + Cfg::STAGE_BYTES);
    uint64_t *x_full = bars, *x_empty = bars + 2, *w_full = bars + 4, *w_empty = bars + 4 + NW;
    uint64_t *acc1_full = bars + 4 + 2 * NW /*[2]: per 128-row half*/, *e1_done = acc1_full + 2 /*[2]*/,
             *acc2_full = acc1_full + 4 /*[2]*/, *acc2_empty = acc1_full + 6 /*[2]*/;
    uint32_t *tmem_slot = (uint32_t *)(acc1_full + 8);
    float *s_se = (float *)(bars + 64);
    float *s_e1par = s_se + Cfg::SE_FLOATS;   // epilogue 1: BN2 scale[128] | BN2 shift + scale * conv1 bias [128]
    volatile int *s_done = (volatile int *)(s_e1par + 2 * 128);   // [8]: items whose bf16 stores epilogue-2 warp w has seen complete

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int pair0 = (int)(blockIdx.x >> 1), pair_step = (int)(gridDim.x >> 1);
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const int n_tiles = (int)(((long long)cnt * p.P_pad + TILE_ROWS - 1) / TILE_ROWS);   // a tile = 256 / P_pad whole boards
    const int bpt = TILE_ROWS / p.P_pad;
    const int n_loop = (n_tiles + 1) / 2;
    const int n_my = n_loop > pair0 ? (n_loop - pair0 + pair_step - 1) / pair_step : 0;   // items of this CTA per layer
    const int dbg = p.dbg;
    const int NL = p.n_layers;

    if (threadIdx.x < 8) s_done[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
        for (int s = 0; s < NW; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int h = 0; h < 2; h++) { mbar_init(&acc1_full[h], 1); mbar_init(&e1_done[h], 8); } // 4 epilogue-1 warps of each CTA
        for (int a = 0; a < 2; a++) { mbar_init(&acc2_full[a], 1); mbar_init(&acc2_empty[a], 16); } // 8 epilogue-2 warps of each CTA
        fence_barrier_init();
        for (int l = 0; l < NL; l++) {
            tma_prefetch_desc(&p.L[l].tmA);
            tma_prefetch_desc(&p.L[l].tmW1);
            tma_prefetch_desc(&p.L[l].tmW2);
            if (p.L[l].out_a) tma_prefetch_desc(&p.L[l].tmOa);
            if (p.L[l].out_b) tma_prefetch_desc(&p.L[l].tmOb);
        }
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The two 128-row halves of a board ("sub" 0 / 1) are scheduled separately at the two hand-overs between the tensor
    // pipe and epilogue 1, so that the pipe keeps working while conv1's accumulator is drained (C_in = 128; `defer`):
    //   conv1     K-block 0: all taps, both halves;  K-block 1: taps 0..3 both halves, taps 4..8 half 0   -> acc1_full[0]
    //          C: K-block 1, taps 4..8 of half 1 (they read slab rows >= 128 only)                        -> acc1_full[1]
    //             ... while epilogue 1 rewrites rows 0..127 of the slab in place (half 0)
    //   conv2  D: K-block 0, taps 0..5 of half 0 (rows <= 128; row 128 only for the dead output row 127)   after e1_done[0]
    //             ... while epilogue 1 rewrites rows 128..255 (half 1)
    //          E: K-block 0, taps 0..5 of half 1, taps 6..8 of both; K-block 1 whole                       after e1_done[1]
    // The weight tiles of C and of D's second visit are still in the ring: the issuer keeps them (no commit to w_empty) on
    // the first visit and rewinds its ring position for the second, so the producer's order and the 36 tiles per board
    // are unchanged, and so is the order in which every accumulator row sees its (K-block, tap) terms.
    // C_in = 256 (K-blocks 2, 3 overwrite 0, 1): conv1 is issued whole, conv2 as above.
    const uint32_t sH_addr = smem_u32(sH), par_addr = smem_u32(s_e1par);
    const bool e1_off = (dbg & 512) != 0;

    if (warp == 0) { // ---------------- slab TMA producer (one lane)
        if (lane == 0) {
            uint32_t fills[2] = {0u, 0u};     // how often each slab has been filled
            int it = 0;
            for (int l = 0; l < NL; l++) {
                const TrunkLayer &L = p.L[l];
                for (int lt = pair0; lt < n_loop; lt += pair_step, it++) { // K-block kc of the input goes to slab kc & 1
                    const int t = 2 * lt + rank;
                    const int row0 = t * TILE_ROWS - HALO;
                    if (L.dep) {   // the previous layer's outputs for this tile (item it - n_my) have reached global memory
                        const int need = it - n_my + 1;
                        for (int w = 0; w < 8; w++)
                            while (s_done[w] < need) { }
                        __threadfence();
                    }
                    for (int kc = 0; kc < L.nkc1; kc++) {
                        const int sl = kc & 1;
                        mbar_wait(&x_empty[sl], (fills[sl] & 1) ^ 1);
                        fills[sl]++;
                        if (rank == 0) mbar_expect_tx(&x_full[sl], 2 * SLAB_BYTES);
                        uint8_t *dst = sX + sl * SLAB_BYTES;
                        tma_load_2d_pair(dst, &L.tmA, &x_full[sl], kc * 64, row0);
                        tma_load_2d_pair(dst + SLAB_BYTES / 2, &L.tmA, &x_full[sl], kc * 64, row0 + SLAB_BOX_ROWS);
                    }
                }
            }
        }
    } else if (warp == 1) { // ---------------- weight-tile TMA producer (one lane), tiles in the issuer's order
        if (lane == 0) {
            Ring r;
            auto load = [&](const CUtensorMap *tm, int kc, int tap, int cin) {
                mbar_wait(&w_empty[r.idx], r.phase ^ 1);
                if (rank == 0) mbar_expect_tx(&w_full[r.idx], 2 * Cfg::W_BYTES);
                tma_load_2d_pair(sW + r.idx * Cfg::W_BYTES, tm, &w_full[r.idx], tap * cin + kc * 64, rank * 64);
                r.advance(NW);
            };
            for (int l = 0; l < NL; l++) {
                const TrunkLayer &L = p.L[l];
                const int cin1 = L.nkc1 * 64;
                for (int lt = pair0; lt < n_loop; lt += pair_step) {
                    for (int kc = 0; kc < L.nkc1; kc++)
                        for (int tap = 0; tap < 9; tap++) load(&L.tmW1, kc, tap, cin1);
                    for (int kc = 0; kc < 2; kc++)
                        for (int tap = 0; tap < 9; tap++) load(&L.tmW2, kc, tap, 128);
                }
            }
        }
    } else if (warp == 2) {
        if (rank == 0) { // ---------------- MMA issuer (leader CTA): whole warp, one elected lane per instruction
            constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
            Ring rw;
            uint32_t ph = 0;
            uint32_t got[2] = {0u, 0u};       // how often each slab has been consumed as a conv1 input
            int it = 0;
            [[maybe_unused]] long long tk_e2 = 0, tk_e1 = 0, tk_w = 0, tk_x = 0, tk0 = 0, tq = 0;
            TK_START(tk0);
            const uint32_t slab_lo0 = umma_desc_lo(smem_u32(sX) + (uint32_t)(HALO * 128)),
                           slab_lo1 = umma_desc_lo(smem_u32(sX + SLAB_BYTES) + (uint32_t)(HALO * 128));
#define TILE(SUBS_, sl_, kc_, tap_, RELEASE_) do {                                                                            \
        TK_BEGIN();                                                                                                    \
        mbar_wait(&w_full[rw.idx], rw.phase);                                                                          \
        TK_END(tk_w);                                                                                                  \
        tc_fence_after();                                                                                              \
        issue_tile<SUBS_>(d0, (sl_) ? slab_lo1 : slab_lo0, umma_desc_lo(smem_u32(sW + rw.idx * Cfg::W_BYTES)), tap_, p.Wp,            \
                          (uint32_t)(((kc_) | (tap_)) != 0), idesc);                                                   \
        if (RELEASE_) umma_commit_elect<true>(&w_empty[rw.idx]);                                                       \
        rw.advance(NW);                                                                                                \
    } while (0)
            for (int l = 0; l < NL; l++) {
                const int nkc1 = p.L[l].nkc1;
                const bool defer = nkc1 == 2;
                for (int lt = pair0; lt < n_loop; lt += pair_step, ph ^= 1, it++) {
                    // Accumulator set `as` (256 TMEM columns) serves BOTH convolutions of this board: conv1 fills it, epilogue 1
                    // drains it into slab H, conv2 refills it, epilogue 2 reads it - while the next board already runs both of
                    // its convolutions in the other set.
                    const int as = it & 1;
                    const uint32_t sph = (uint32_t)((it >> 1) & 1);
                    const uint32_t d0 = tmem_base + (uint32_t)(as * 2 * BN);
                    // ---- conv1
                    TK_BEGIN();
                    mbar_wait(&acc2_empty[as], sph ^ 1);  // epilogue 2 of board it-2 has drained this set
                    TK_END(tk_e2);
                    tc_fence_after();
                    Ring hold;
                    for (int kc = 0; kc < nkc1; kc++) {
                        const int sl = kc & 1;
                        TK_BEGIN();
                        mbar_wait(&x_full[sl], got[sl] & 1u);
                        got[sl]++;
                        TK_END(tk_x);
                        tc_fence_after();
                        for (int tap = 0; tap < 4; tap++) TILE(3, sl, kc, tap, true);
                        if (defer && kc == 1) {
                            hold = rw;
                            for (int tap = 4; tap < 9; tap++) TILE(1, sl, kc, tap, false);   // tiles stay for phase C
                        } else {
                            for (int tap = 4; tap < 9; tap++) TILE(3, sl, kc, tap, true);
                        }
                        if (kc + 2 < nkc1) umma_commit_elect<true>(&x_empty[sl]);   // C_in = 256: the slab takes K-block kc + 2 next
                    }
                    umma_commit_elect<true>(&acc1_full[0]);
                    if (defer) {
                        rw = hold;
                        for (int tap = 4; tap < 9; tap++) TILE(2, 1, 1, tap, true);          // phase C
                    }
                    umma_commit_elect<true>(&acc1_full[1]);
                    // ---- conv2 (input: slab H = the two slabs, rewritten in place by epilogue 1)
                    TK_BEGIN();
                    mbar_wait(&e1_done[0], ph);            // half 0 of the set drained, rows 0..127 of H written (both CTAs)
                    TK_END(tk_e1);
                    tc_fence_after();
                    hold = rw;
                    for (int tap = 0; tap < 6; tap++) TILE(1, 0, 0, tap, false);             // phase D, tiles stay for E
                    TK_BEGIN();
                    mbar_wait(&e1_done[1], ph);            // half 1 drained, rows 128..255 written
                    TK_END(tk_e1);
                    tc_fence_after();
                    rw = hold;
                    for (int tap = 0; tap < 6; tap++) TILE(2, 0, 0, tap, true);              // phase E
                    for (int tap = 6; tap < 9; tap++) TILE(3, 0, 0, tap, true);
                    umma_commit_elect<true>(&x_empty[0]);  // conv2 has read K-block 0 of H: the slab is free for the next board
                    for (int tap = 0; tap < 9; tap++) TILE(3, 1, 1, tap, true);
                    umma_commit_elect<true>(&x_empty[1]);
                    umma_commit_elect<true>(&acc2_full[as]);
                }
            }
#undef TILE
            if ((dbg & 1024) && blockIdx.x == 0 && lane == 0)
            {
                TK_STOP(tk0);
                TK_PRINT("issuer: boards %d total %lld wait_e2 %lld wait_e1 %lld wait_weights %lld wait_slab %lld\n", it, tk0, tk_e2, tk_e1,
                         tk_w, tk_x);
            }
        }
    } else if (warp >= 4 && warp < 8) { // ---------------- epilogue 1 (acc1 -> relu(BN2(conv1 + b1)) -> slab, bf16, swizzled), half 0 then half 1
        const int e1_q = warp & 3;
        const int e1t = threadIdx.x - 128;   // 0..127
        int it = 0;
        [[maybe_unused]] long long tk_work = 0, tq = 0;
        for (int l = 0; l < NL; l++) {
            {   // this layer's BN2 scale and bias-folded shift (the four epilogue-1 warps own the table)
                named_bar_sync(2, 128);
                const float sc = p.L[l].par1[128 + e1t];
                s_e1par[e1t] = sc;
                s_e1par[128 + e1t] = fmaf(sc, p.L[l].par1[e1t], p.L[l].par1[256 + e1t]);
                named_bar_sync(2, 128);
            }
            for (int lt = pair0; lt < n_loop; lt += pair_step, it++) {
                const bool work = (2 * lt + rank) < n_tiles && !e1_off;
                TK_BEGIN();
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    const int pos = h * 128 + e1_q * 32 + lane;
                    const int bp = pos % p.P_pad, yy = bp / p.Wp;
                    const bool live = yy != 0 && yy <= p.H && (bp % p.Wp) != p.Wp - 1 && (2 * lt + rank) * bpt + pos / p.P_pad < cnt;
                    const uint32_t t_sub = tmem_base + ((uint32_t)(e1_q * 32) << 16) + (uint32_t)((it & 1) * 2 * BN + h * BN);
                    e1_half(&acc1_full[h], &e1_done[h], (uint32_t)(it & 1), work, t_sub, par_addr, sH_addr, HALO + pos, live, lane);
                }
                TK_END(tk_work);
            }
        }
        if ((dbg & 1024) && blockIdx.x == 0 && warp == 4 && lane == 0) TK_PRINT("epilogue1: wait + work %lld\n", tk_work);
    } else if (warp >= 8) { // ---------------- epilogue 2: SE + skip add + outputs (acc2)
        const int ew = warp - 8;
        const int q = warp & 3, sub = ew >> 2;
        const int et = threadIdx.x - 256; // 0..255
        // s_se: column-sum partials [8 warps][128] (later the two dense2 partials) | mean[128] | dense1 partials [4][64] |
        // hidden[64] | gate[128] | gate * conv2 bias [128]
        float *s_part = s_se, *s_mean = s_se + 8 * BN, *s_hp = s_mean + BN, *s_h = s_hp + 256, *s_gate = s_h + 64, *s_bg = s_gate + BN;
        float *s_gp = s_part;
        const float inv_cells = 1.0f / (float)p.n_cells;
        const int pos = sub * 128 + q * 32 + lane;          // row of the tile: the same for every tile of this thread
        const int bp = pos % p.P_pad;
        const bool live_pos = (bp / p.Wp) != 0 && (bp / p.Wp) <= p.H && (bp % p.Wp) != p.Wp - 1;
        const uint32_t gate_addr = smem_u32(s_gate), bg_addr = smem_u32(s_bg);
        const uint32_t stage_addr = smem_u32(sStage + ew * 2 * 2048);
        int it = 0;
        [[maybe_unused]] long long tk_wait = 0, tk_se = 0, tk_out = 0, tq = 0;
        for (int l = 0; l < NL; l++) {
            const TrunkLayer &L = p.L[l];
            const bool use_res = L.res && !(dbg & 8);
            const bool do_se = L.se && !(dbg & 64);
            if (!do_se) { // no gate: out = conv2 + bias (+ residual); the table is rewritten once every warp has left the previous layer
                named_bar_sync(1, 256);
                if (et < BN) { s_gate[et] = 1.0f; s_bg[et] = L.par2[et]; }
                named_bar_sync(1, 256);
            }
            for (int lt = pair0; lt < n_loop; lt += pair_step, it++) {
                const int t = 2 * lt + rank;
                const int as = it & 1;
                const uint32_t sph = (uint32_t)((it >> 1) & 1);
                if (t >= n_tiles) { // dummy half of the last pair
                    mbar_wait(&acc2_full[as], sph);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive_leader(&acc2_empty[as]); s_done[ew] = it + 1; }
                    continue;
                }
                // blocked fp32 layout (f32_blk_index): the 32 rows x 128 channels of this warp are 16 KB contiguous; 8-channel
                // piece k of this thread's row sits at + k * 256 floats
                const size_t rbase = ((size_t)(t * 8 + sub * 4 + q) << 12) + (size_t)(lane * 8);
                const float *resp = L.res + rbase;
                float *outp = L.out_raw + rbase;
                const int row0 = t * TILE_ROWS + sub * 128 + q * 32;   // first row of this warp (TMA store coordinate)
                const bool live = live_pos && t * bpt + pos / p.P_pad < cnt;     // boards past the last one of the batch stay zero
                const uint32_t mask = live ? 0xffffffffu : 0u;
                float resA[16], resB[16];
                if (use_res) {
#pragma unroll 1
                    for (int k = 4; k < 16; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(resp + k * 256));
                    ldg256(resp, *reinterpret_cast<float(*)[8]>(&resA[0]));
                    ldg256(resp + 256, *reinterpret_cast<float(*)[8]>(&resA[8]));
                }
                TK_BEGIN();
                mbar_wait(&acc2_full[as], sph);
                TK_END(tk_wait);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * BN + sub * BN);
                TK_BEGIN();
                if (do_se) {
                    // ---- pass 1: per-channel sums over the live cells of this warp's 32 rows (32-column pieces, the TMEM load
                    // of the next piece in flight while this one is folded)
                    {
                        uint32_t ra[32], rb[32];
                        float *dst = s_part + ew * BN;
                        tmem_ld_32x32(t_acc, ra);
#pragma unroll 1
                        for (int h = 0; h < 2; h++) {
                            tmem_ld_wait_dep(ra);
                            tmem_ld_32x32(t_acc + (uint32_t)(h * 64 + 32), rb);
                            colsum32(ra, mask, lane, dst + h * 64);
                            tmem_ld_wait_dep(rb);
                            if (h == 0) tmem_ld_32x32(t_acc + 64u, ra);
                            colsum32(rb, mask, lane, dst + h * 64 + 32);
                        }
                    }
                    named_bar_sync(1, 256);
                    if (et < BN) { // board mean per channel
                        float sum = 0.0f;
#pragma unroll
                        for (int k = 0; k < 8; k++) sum += s_part[k * BN + et];
                        s_mean[et] = sum * inv_cells;
                    }
                    named_bar_sync(1, 256);
                    {   // dense1 (C -> R): output j, quarter `part` of the inputs
                        const int j = et & 63, part = et >> 6;
                        if (j < L.se_r) {
                            const float *w1 = L.se_w1 + (size_t)(part * 32) * L.se_r + j;
                            const float4 *m4 = reinterpret_cast<const float4 *>(s_mean + part * 32);
                            float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const float4 m = m4[i];
                                h0 = fmaf(m.x, __ldg(w1 + (size_t)(4 * i) * L.se_r), h0);
                                h1 = fmaf(m.y, __ldg(w1 + (size_t)(4 * i + 1) * L.se_r), h1);
                                h2 = fmaf(m.z, __ldg(w1 + (size_t)(4 * i + 2) * L.se_r), h2);
                                h3 = fmaf(m.w, __ldg(w1 + (size_t)(4 * i + 3) * L.se_r), h3);
                            }
                            s_hp[part * 64 + j] = (h0 + h1) + (h2 + h3);
                        }
                    }
                    named_bar_sync(1, 256);
                    if (et < L.se_r) s_h[et] = fmaxf(((s_hp[et] + s_hp[64 + et]) + (s_hp[128 + et] + s_hp[192 + et])) + __ldg(L.se_b1 + et), 0.0f);
                    named_bar_sync(1, 256);
                    {   // dense2 (R -> C): output channel cc, half `part` of the hidden units
                        const int cc = et & 127, part = et >> 7;
                        const int r2 = L.se_r >> 1;
                        const float *w2 = L.se_w2 + (size_t)(part * r2) * BN + cc;
                        const float *hh = s_h + part * r2;
                        float g0 = 0.0f, g1 = 0.0f;
#pragma unroll 8
                        for (int i = 0; i < r2; i += 2) {
                            g0 = fmaf(hh[i], __ldg(w2 + (size_t)(i) * BN), g0);
                            g1 = fmaf(hh[i + 1], __ldg(w2 + (size_t)(i + 1) * BN), g1);
                        }
                        s_gp[part * BN + cc] = g0 + g1;
                    }
                    named_bar_sync(1, 256);
                    if (et < BN) {
                        const float g = 1.0f / (1.0f + expf(-((s_gp[et] + s_gp[BN + et]) + L.se_b2[et])));
                        s_gate[et] = g;
                        s_bg[et] = g * L.par2[et];   // gate * (conv2 + bias) = fma(conv2, gate, gate * bias)
                    }
                    named_bar_sync(1, 256);
                }
                TK_END(tk_se);
                TK_BEGIN();
                // ---- output pass: 16-column chunks in pairs (A, B); TMEM load and residual load of the next chunk are in
                // flight while the current one is written
                {
                    uint32_t accA[16], accB[16];
                    tmem_ld_32x16(t_acc, accA);
#pragma unroll 1
                    for (int cp = 0; cp < 4; cp++) {
                        const int ck = 2 * cp;
                        tmem_ld_wait_dep(accA);
                        tmem_ld_32x16(t_acc + (uint32_t)(ck * 16 + 16), accB);
                        if (use_res) {
                            ldg256(resp + (2 * ck + 2) * 256, *reinterpret_cast<float(*)[8]>(&resB[0]));
                            ldg256(resp + (2 * ck + 3) * 256, *reinterpret_cast<float(*)[8]>(&resB[8]));
                        }
                        e2_chunk<0>(accA, resA, use_res, ck, L, outp, gate_addr, bg_addr, stage_addr, mask, lane, row0, dbg);
                        tmem_ld_wait_dep(accB);
                        if (cp < 3) {
                            tmem_ld_32x16(t_acc + (uint32_t)(ck * 16 + 32), accA);
                            if (use_res) {
                                ldg256(resp + (2 * ck + 4) * 256, *reinterpret_cast<float(*)[8]>(&resA[0]));
                                ldg256(resp + (2 * ck + 5) * 256, *reinterpret_cast<float(*)[8]>(&resA[8]));
                            }
                        } else { // the accumulator set is in registers: hand it back to the MMA issuer before the last stores
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_leader(&acc2_empty[as]);
                        }
                        e2_chunk<1>(accB, resB, use_res, ck + 1, L, outp, gate_addr, bg_addr, stage_addr, mask, lane, row0, dbg);
                    }
                }
                // a following layer of this launch reads this tile's bf16 operand back through TMA: publish once the stores of
                // this warp have completed (the fp32 stream is re-read by the very thread that wrote it)
                if (l + 1 < NL) {
                    if (lane == 0) {
                        tma_store_wait_all();
                        __threadfence();
                        s_done[ew] = it + 1;
                    }
                    __syncwarp();
                }
                TK_END(tk_out);
            }
        }
        if ((dbg & 1024) && blockIdx.x == 0 && (warp == 8 || warp == 15) && lane == 0)
            TK_PRINT("epilogue2 warp %d: wait_acc %lld se %lld out %lld\n", warp, tk_wait, tk_se, tk_out);
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

} // namespace gaz_block
