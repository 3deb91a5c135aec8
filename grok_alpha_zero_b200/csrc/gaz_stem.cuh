// gaz_stem.cuh -- the network stem (Net/ResNet/*: first 3x3 convolution on the 2 input planes + BN [+ ReLU]) on the
// tensor cores for tile == board geometries (Gomoku: 256 padded rows per board, 256 stem filters).
//
// The CUDA-core stem (stem_kernel in gaz_net.cu) is instruction-bound: 2 x 256 channels x 256 rows of bf16 outputs per
// board cost ~50k warp instructions.  Here a board is one small implicit GEMM: A = im2col of the board built in shared
// memory (256 rows x K, K = 9 taps x 2 planes; the entries are -1/0/+1, exact in bf16), B = the filters, D in TMEM, and
// the only real work left is the store-bound epilogue (BN affine, optional ReLU, the two bf16 operands of the first
// residual block through swizzled staging tiles + TMA stores).
// fp32 filter accuracy is kept by splitting every weight into bf16 hi + lo parts along K: k = 0..17 multiplies the hi
// parts, k = 18..35 the lo parts (A repeats its 18 entries), K is padded to 48 = three k-steps of 16.
// One CTA computes HALF of the output channels (N = 128) of one board at a time: 256 TMEM columns (2 x 128-row halves),
// so two CTAs share an SM and one CTA's epilogue overlaps the other's im2col + MMA without any pipeline inside a CTA.
#pragma once
#include "gaz_conv.cuh"
#include "gaz_small.cuh"

namespace gaz_stem {
using namespace gaz_tc;

struct StemTcArgs {
    const int32_t *count;
    int max_count;
    const int8_t *states;   // [leaf][H*W*2] HWC (plane 0 = side to move, plane 1 = stones), values -1/0/+1
    int H, W, Wp, relu;
    const uint16_t *wpack;  // [256 filters][64 k] bf16: k = tap*2 + plane (hi parts) | 18 + tap*2 + plane (lo parts) | zeros
    const float *par;       // [4][256]: BN scale | BN shift + scale * conv bias | scale_a | shift_a
    int has_q, has_a;       // out_q = activation itself (bf16), out_a = relu(scale_a * activation + shift_a) (bf16)
};

struct Cfg {
    static constexpr int A_BYTES = 256 * 128, B_BYTES = 128 * 128, STAGE_BYTES = 8 * 2 * 2048, PAR_BYTES = 4 * 128 * 4;
    static constexpr int IN_BYTES = 18 * 18 * 2 + 28;  // zero-bordered int8 board (H, W <= 16)
    static constexpr int SMEM = 1024 + A_BYTES + B_BYTES + STAGE_BYTES + PAR_BYTES + IN_BYTES + 64;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// one 32-column piece of one row: BN affine (+ReLU) -> out_q / out_a staging tiles (32 rows x 32 channels, SWIZZLE_64B)
__device__ __forceinline__ void stem_piece(const uint32_t (&r)[32], uint32_t par_addr, int c32, bool relu, bool has_q, bool has_a,
                                           uint32_t stage_addr, uint32_t mask, int lane) {
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
    for (int ch = 0; ch < 4; ch++) { // 8 channels = one 16-byte chunk of each output tile
        const uint32_t o = (uint32_t)(c32 * 128 + ch * 32);
        const float4 s0 = lds128f(par_addr + o), s1 = lds128f(par_addr + o + 16);
        const float4 h0 = lds128f(par_addr + 512 + o), h1 = lds128f(par_addr + 512 + o + 16);
        float v[8];
        const int j = ch * 8;
        v[0] = fmaf(s0.x, __uint_as_float(r[j]), h0.x);     v[1] = fmaf(s0.y, __uint_as_float(r[j + 1]), h0.y);
        v[2] = fmaf(s0.z, __uint_as_float(r[j + 2]), h0.z); v[3] = fmaf(s0.w, __uint_as_float(r[j + 3]), h0.w);
        v[4] = fmaf(s1.x, __uint_as_float(r[j + 4]), h1.x); v[5] = fmaf(s1.y, __uint_as_float(r[j + 5]), h1.y);
        v[6] = fmaf(s1.z, __uint_as_float(r[j + 6]), h1.z); v[7] = fmaf(s1.w, __uint_as_float(r[j + 7]), h1.w);
        if (relu) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = fmaxf(v[i], 0.0f);
        }
        const uint32_t off = (uint32_t)(lane * 64) + (((uint32_t)ch ^ sw) << 4);
        if (has_q)
            sts128(stage_addr + off, pack_bf16x2(v[0], v[1]) & mask, pack_bf16x2(v[2], v[3]) & mask, pack_bf16x2(v[4], v[5]) & mask,
                   pack_bf16x2(v[6], v[7]) & mask);
        if (has_a) {
            const float4 a0 = lds128f(par_addr + 1024 + o), a1 = lds128f(par_addr + 1024 + o + 16);
            const float4 t0 = lds128f(par_addr + 1536 + o), t1 = lds128f(par_addr + 1536 + o + 16);
            sts128(stage_addr + 2048 + off, pack_relu_bf16x2(fmaf(a0.x, v[0], t0.x), fmaf(a0.y, v[1], t0.y)) & mask,
                   pack_relu_bf16x2(fmaf(a0.z, v[2], t0.z), fmaf(a0.w, v[3], t0.w)) & mask,
                   pack_relu_bf16x2(fmaf(a1.x, v[4], t1.x), fmaf(a1.y, v[5], t1.y)) & mask,
                   pack_relu_bf16x2(fmaf(a1.z, v[6], t1.z), fmaf(a1.w, v[7], t1.w)) & mask);
        }
    }
}

__global__ void __launch_bounds__(256, 2)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmOq, const __grid_constant__ CUtensorMap tmOa, const StemTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base, *sB = base + Cfg::A_BYTES, *sStage = sB + Cfg::B_BYTES;
    float *s_par = (float *)(sStage + Cfg::STAGE_BYTES);
    int8_t *s_in = (int8_t *)(s_par + 4 * 128);
    uint64_t *bar = (uint64_t *)(((uintptr_t)(s_in + Cfg::IN_BYTES) + 7) & ~(uintptr_t)7);
    uint32_t *tmem_slot = (uint32_t *)(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nhalf = (int)(blockIdx.x & 1);              // which 128 of the 256 filters
    const int b0 = (int)(blockIdx.x >> 1), bstep = (int)(gridDim.x >> 1);
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;

    // one-time setup: filters (swizzled like a SWIZZLE_128B TMA write), parameters, zeroed A tile and input border
    for (int i = threadIdx.x; i < 128 * 8; i += 256) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4 *>(sB + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4 *>(p.wpack + (size_t)(nhalf * 128 + r) * 64 + c * 8);
    }
    for (int i = threadIdx.x; i < 256 * 8; i += 256) *reinterpret_cast<uint4 *>(sA + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < 4 * 128; i += 256) s_par[i] = p.par[(i >> 7) * 256 + nhalf * 128 + (i & 127)];
    for (int i = threadIdx.x; i < Cfg::IN_BYTES; i += 256) s_in[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        if (p.has_q) tma_prefetch_desc(&tmOq);
        if (p.has_a) tma_prefetch_desc(&tmOa);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int pos = threadIdx.x;                           // padded row of the board owned by this thread (im2col + epilogue)
    const int yy = pos / p.Wp - 1, xx = pos % p.Wp;
    const bool live = yy >= 0 && yy < p.H && xx < p.W;
    const uint32_t mask = live ? 0xffffffffu : 0u;
    const int pitch = (p.W + 2) * 2;                       // zero-bordered board, 2 bytes (planes) per cell
    const int q = warp & 3, sub = warp >> 2;               // TMEM lane quarter / 128-row half: rows sub*128 + q*32 + lane = pos
    const uint32_t a_row = smem_u32(sA) + (uint32_t)(pos * 128);
    const uint32_t par_addr = smem_u32(s_par), stage_addr = smem_u32(sStage + warp * 2 * 2048);
    const uint32_t a_lo = umma_desc_lo(smem_u32(sA)), b_lo = umma_desc_lo(smem_u32(sB));
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    const int ncell = p.H * p.W;
    uint32_t ph = 0;

    for (int b = b0; b < cnt; b += bstep, ph ^= 1) {
        // ---- board -> zero-bordered int8 tile
        if (threadIdx.x < ncell) {
            const int y = threadIdx.x / p.W, x = threadIdx.x - y * p.W;
            *reinterpret_cast<uint16_t *>(s_in + (y + 1) * pitch + (x + 1) * 2) =
                *reinterpret_cast<const uint16_t *>(p.states + ((size_t)b * ncell + threadIdx.x) * 2);
        }
        __syncthreads();
        // ---- im2col row: 9 taps x 2 planes (k = tap*2 + plane), repeated for the lo parts of the weights
        if (live) {
            uint32_t w[10];
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const uint16_t two = *reinterpret_cast<const uint16_t *>(s_in + (yy + t / 3) * pitch + (xx + t % 3) * 2);
                const int v0 = (int)(int8_t)(two & 0xff), v1 = (int)(int8_t)(two >> 8);
                // -1 / 0 / +1 -> bf16 bits 0xBF80 / 0 / 0x3F80
                const uint32_t h0 = v0 == 0 ? 0u : (v0 > 0 ? 0x3F80u : 0xBF80u), h1 = v1 == 0 ? 0u : (v1 > 0 ? 0x3F80u : 0xBF80u);
                w[t] = h0 | (h1 << 16);
            }
            w[9] = 0u;
            const uint32_t sw = (uint32_t)(pos & 7);
            // words 0..8 = k 0..17, words 9..17 = k 18..35 (same entries), words 18, 19 = 0
            sts128(a_row + ((0u ^ sw) << 4), w[0], w[1], w[2], w[3]);
            sts128(a_row + ((1u ^ sw) << 4), w[4], w[5], w[6], w[7]);
            sts128(a_row + ((2u ^ sw) << 4), w[8], w[0], w[1], w[2]);
            sts128(a_row + ((3u ^ sw) << 4), w[3], w[4], w[5], w[6]);
            sts128(a_row + ((4u ^ sw) << 4), w[7], w[8], w[9], w[9]);
        }
        fence_proxy_async();
        __syncthreads();
        // ---- D[256 rows][128 filters] = A[256][48] x B[128][48]^T : 2 halves x 3 k-steps
        if (warp == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int k = 0; k < 3; k++)
                    umma_bf16_elect<false>(tmem_base + (uint32_t)(s * 128), a_lo + (uint32_t)(s * 1024 + k * 2), b_lo + (uint32_t)(k * 2),
                                           idesc, k > 0 ? 1u : 0u);
            umma_commit_elect<false>(bar);
        }
        mbar_wait(bar, ph);
        tc_fence_after();
        // ---- epilogue: this thread's row, 4 pieces of 32 filters
        {
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 128);
            const int row0 = b * 256 + sub * 128 + q * 32;
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(t_acc, ra);
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                tmem_ld_wait_dep(ra);
                tmem_ld_32x32(t_acc + (uint32_t)(c * 64 + 32), rb);
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                stem_piece(ra, par_addr, 2 * c, p.relu != 0, p.has_q != 0, p.has_a != 0, stage_addr, mask, lane);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (p.has_q) tma_store_2d_addr(&tmOq, stage_addr, nhalf * 128 + c * 64, row0);
                    if (p.has_a) tma_store_2d_addr(&tmOa, stage_addr + 2048, nhalf * 128 + c * 64, row0);
                    tma_store_commit();
                }
                tmem_ld_wait_dep(rb);
                if (c == 0) tmem_ld_32x32(t_acc + 64u, ra);
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                stem_piece(rb, par_addr, 2 * c + 1, p.relu != 0, p.has_q != 0, p.has_a != 0, stage_addr, mask, lane);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (p.has_q) tma_store_2d_addr(&tmOq, stage_addr, nhalf * 128 + c * 64 + 32, row0);
                    if (p.has_a) tma_store_2d_addr(&tmOa, stage_addr + 2048, nhalf * 128 + c * 64 + 32, row0);
                    tma_store_commit();
                }
            }
        }
        tc_fence_before();
        __syncthreads();   // TMEM, the A tile and the input tile are free for the next board
    }
    if (lane == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---- stem + the 1x1 projection of the first residual block's shortcut in one kernel ------------------------------------------
// Net/ResNet/ResNet_Block.py:29-31: when the stem has more filters than the trunk (Gomoku: 256 -> 128) the first block's
// shortcut is a 1x1 convolution of the stem output.  As two kernels the stem writes that output as a second bf16 operand
// (128 KB per board) only for the projection to read it back; here it stays in shared memory: per 128-row half of a board
//   im2col (K = 48) --MMA N=256--> TMEM --epilogue 1: BN (+ReLU)--> x0 as bf16 K-major SWIZZLE_128B tiles in shared memory
//                                        + relu(BN1(x0)) --staging tiles, TMA store--> out_a (operand of the block's conv1)
//   x0 tiles --MMA N=128, K=256 (projection filters resident in shared memory)--> TMEM --epilogue 2: + bias--> fp32 stream
// HBM traffic per board: 128 KB (out_a) + 128 KB (fp32 shortcut) instead of 512 KB.  Same operand values and the same
// accumulation order as stem_tc_kernel followed by conv_board_kernel<128> (1x1), so the results are bit-identical.
// 512 threads: warp w works on TMEM lane quarter w & 3 and column part w >> 2 (epilogue 1: 64 of the 256 stem filters =
// one 64-channel K-block of the projection's A operand; epilogue 2: 32 of the 128 outputs).
struct StemProjArgs {
    const int32_t *count;
    int max_count;
    const int8_t *states;
    int H, W, Wp, relu;
    const uint16_t *wpack;  // as StemTcArgs
    const float *par;       // as StemTcArgs
    const uint16_t *wproj;  // [128 outputs][256 k] bf16, K-major (the 1x1 convolution's tensor-core weights)
    const float *pbias;     // [128]
    float *out_res;         // blocked fp32 rows x 128 (gaz_conv::f32_blk_index)
};

struct ProjCfg {
    static constexpr int A1_BYTES = 128 * 128, B1_BYTES = 256 * 128, A2_BYTES = 4 * 128 * 128, B2_BYTES = 4 * 128 * 128;
    static constexpr int STAGE_BYTES = 16 * 2048, PAR_BYTES = 4 * 256 * 4 + 128 * 4;
    static constexpr int IN_BYTES = 18 * 18 * 2 + 28;
    static constexpr int SMEM = 1024 + A1_BYTES + B1_BYTES + A2_BYTES + B2_BYTES + STAGE_BYTES + PAR_BYTES + 2 * IN_BYTES + 64;
};

__global__ void __launch_bounds__(512, 1)
stem_proj_kernel(const __grid_constant__ CUtensorMap tmOa, const StemProjArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA1 = base, *sB1 = sA1 + ProjCfg::A1_BYTES, *sA2 = sB1 + ProjCfg::B1_BYTES, *sB2 = sA2 + ProjCfg::A2_BYTES;
    uint8_t *sStage = sB2 + ProjCfg::B2_BYTES;
    float *s_par = (float *)(sStage + ProjCfg::STAGE_BYTES);   // [4][256] | projection bias [128]
    int8_t *s_in = (int8_t *)(s_par + 4 * 256 + 128);          // two zero-bordered boards (double buffer)
    uint64_t *bars = (uint64_t *)(((uintptr_t)(s_in + 2 * ProjCfg::IN_BYTES) + 7) & ~(uintptr_t)7);
    uint32_t *tmem_slot = (uint32_t *)(bars + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;

    for (int i = threadIdx.x; i < 256 * 8; i += 512) {       // stem filters, swizzled like a SWIZZLE_128B TMA write
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4 *>(sB1 + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(p.wpack + (size_t)r * 64 + c * 8);
    }
    for (int i = threadIdx.x; i < 128 * 32; i += 512) {      // projection filters: four 64-channel K-blocks of [128][64]
        const int r = i >> 5, cc = i & 31, kb = cc >> 3, c = cc & 7;
        *reinterpret_cast<uint4 *>(sB2 + kb * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4 *>(p.wproj + (size_t)r * 256 + cc * 8);
    }
    for (int i = threadIdx.x; i < 4 * 256; i += 512) s_par[i] = p.par[i];
    if (threadIdx.x < 128) s_par[4 * 256 + threadIdx.x] = p.pbias[threadIdx.x];
    for (int i = threadIdx.x; i < 2 * ProjCfg::IN_BYTES; i += 512) s_in[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmOa);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int q = warp & 3, part = warp >> 2;
    const int rsub = q * 32 + lane;                          // row of the 128-row half owned by this thread in both epilogues
    const int pitch = (p.W + 2) * 2;
    const int ncell = p.H * p.W;
    const uint32_t par_addr = smem_u32(s_par), stage_addr = smem_u32(sStage + warp * 2048);
    const uint32_t a1_lo = umma_desc_lo(smem_u32(sA1)), b1_lo = umma_desc_lo(smem_u32(sB1));
    const uint32_t a2_lo = umma_desc_lo(smem_u32(sA2)), b2_lo = umma_desc_lo(smem_u32(sB2));
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 256), idesc2 = umma_idesc_bf16(128, 128);
    const uint32_t a2_row = smem_u32(sA2) + (uint32_t)(part * 16384 + rsub * 128);
    const uint32_t sw7 = (uint32_t)(rsub & 7), sw3 = (uint32_t)((lane >> 1) & 3);

    // The work list of this CTA: halves i = 0 .. T-1, half i = 128-row half i & 1 of board blockIdx.x + (i >> 1) * gridDim.x.
    // Software pipeline over the halves (one __syncthreads per half):
    //   iteration i:  wait stem(i) | TMEM -> registers | wait projection(i-1) | epilogue 1 of i (x0 tiles, out_a) |
    //                 epilogue 2 of i-1 (fp32 shortcut) | warps 12..15: im2col of i+1 | sync | issue stem(i+1), projection(i)
    // so both MMAs and the im2col run in the shadow of the two epilogues.
    const int n_my = cnt > (int)blockIdx.x ? (cnt - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int T = 2 * n_my;
    auto board_of = [&](int i) { return (int)blockIdx.x + (i >> 1) * (int)gridDim.x; };
    bool live_sub[2];
    for (int h = 0; h < 2; h++) {
        const int pos = h * 128 + rsub;
        const int yy = pos / p.Wp - 1, xx = pos % p.Wp;
        live_sub[h] = yy >= 0 && yy < p.H && xx < p.W;
    }
    auto live_of = [&](int i) { return (i & 1) ? live_sub[1] : live_sub[0]; };
    auto load_states = [&](int k) {     // board k of this CTA's list -> zero-bordered int8 tile k & 1
        if ((int)threadIdx.x < ncell) {
            const int y = threadIdx.x / p.W, x = threadIdx.x - y * p.W;
            *reinterpret_cast<uint16_t *>(s_in + (k & 1) * ProjCfg::IN_BYTES + (y + 1) * pitch + (x + 1) * 2) =
                *reinterpret_cast<const uint16_t *>(p.states + ((size_t)board_of(2 * k) * ncell + threadIdx.x) * 2);
        }
    };
    auto im2col = [&](int i) {          // warps 12..15: the 128 rows of half i, as in stem_tc_kernel; dead rows are written as zeros
        if (threadIdx.x < 384) return;
        const int row = (int)threadIdx.x - 384;
        const int pos = (i & 1) * 128 + row;
        const int yy = pos / p.Wp - 1, xx = pos % p.Wp;
        const bool lv = yy >= 0 && yy < p.H && xx < p.W;
        const int8_t *tile = s_in + ((i >> 1) & 1) * ProjCfg::IN_BYTES;
        uint32_t w[10];
#pragma unroll
        for (int t = 0; t < 9; t++) {
            uint32_t v = 0u;
            if (lv) {
                const uint16_t two = *reinterpret_cast<const uint16_t *>(tile + (yy + t / 3) * pitch + (xx + t % 3) * 2);
                const int v0 = (int)(int8_t)(two & 0xff), v1 = (int)(int8_t)(two >> 8);
                // -1 / 0 / +1 -> bf16 bits 0xBF80 / 0 / 0x3F80
                const uint32_t h0 = v0 == 0 ? 0u : (v0 > 0 ? 0x3F80u : 0xBF80u), h1 = v1 == 0 ? 0u : (v1 > 0 ? 0x3F80u : 0xBF80u);
                v = h0 | (h1 << 16);
            }
            w[t] = v;
        }
        w[9] = 0u;
        const uint32_t a_row = smem_u32(sA1) + (uint32_t)(row * 128), sw = (uint32_t)(row & 7);
        // words 0..8 = k 0..17, words 9..17 = k 18..35 (same entries, they meet the lo parts of the filters), k 36..47 = 0
        sts128(a_row + ((0u ^ sw) << 4), w[0], w[1], w[2], w[3]);
        sts128(a_row + ((1u ^ sw) << 4), w[4], w[5], w[6], w[7]);
        sts128(a_row + ((2u ^ sw) << 4), w[8], w[0], w[1], w[2]);
        sts128(a_row + ((3u ^ sw) << 4), w[3], w[4], w[5], w[6]);
        sts128(a_row + ((4u ^ sw) << 4), w[7], w[8], w[9], w[9]);
        sts128(a_row + ((5u ^ sw) << 4), 0u, 0u, 0u, 0u);
    };
    auto issue_stem = [&]() {           // D1[128 rows][256 filters] (TMEM columns 0..255) = A1[128][48] x B1[256][48]^T
#pragma unroll
        for (int k = 0; k < 3; k++)
            umma_bf16_elect<false>(tmem_base, a1_lo + (uint32_t)(k * 2), b1_lo + (uint32_t)(k * 2), idesc1, k > 0 ? 1u : 0u);
        umma_commit_elect<false>(&bars[0]);
    };
    auto epilogue2 = [&](int i) {       // outputs part*32 .. +31 of this thread's row of half i -> blocked fp32 stream (dead rows as zeros)
        const bool live = live_of(i);
        const int row0 = board_of(i) * 256 + (i & 1) * 128 + q * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(part * 32), r);
        tmem_ld_wait_dep(r);
        float *outp = p.out_res + ((((size_t)(row0 >> 5)) * 4 + (size_t)part) << 10) + (size_t)(lane * 8);
        const uint32_t pb = par_addr + 4096 + (uint32_t)(part * 128);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 b0 = lds128f(pb + j * 32), b1 = lds128f(pb + j * 32 + 16);
            float t[8];
            t[0] = __uint_as_float(r[8 * j]) + b0.x;     t[1] = __uint_as_float(r[8 * j + 1]) + b0.y;
            t[2] = __uint_as_float(r[8 * j + 2]) + b0.z; t[3] = __uint_as_float(r[8 * j + 3]) + b0.w;
            t[4] = __uint_as_float(r[8 * j + 4]) + b1.x; t[5] = __uint_as_float(r[8 * j + 5]) + b1.y;
            t[6] = __uint_as_float(r[8 * j + 6]) + b1.z; t[7] = __uint_as_float(r[8 * j + 7]) + b1.w;
#pragma unroll
            for (int u = 0; u < 8; u++) t[u] = live ? t[u] : 0.0f;
            gaz_conv::stg256(outp + j * 256, t);
        }
    };

    if (T > 0) {
        load_states(0);
        __syncthreads();
        im2col(0);
        if (n_my > 1) load_states(1);
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) { tc_fence_after(); issue_stem(); }
    }
#pragma unroll 1
    for (int i = 0; i < T; i++) {
        const bool live = live_of(i);
        const uint32_t mask = live ? 0xffffffffu : 0u;
        const int row0 = board_of(i) * 256 + (i & 1) * 128 + q * 32;
        // board k + 1 arrives during half 2k (its tile was last read by the im2col of board k - 1, two syncs ago); the global
        // load's latency hides behind the wait for the stem MMA
        if ((i & 1) == 0 && i > 0 && (i >> 1) + 1 < n_my) load_states((i >> 1) + 1);
        mbar_wait(&bars[0], (uint32_t)(i & 1));
        tc_fence_after();
        // ---- epilogue 1: filters part*64 .. +63 of this thread's row
        {
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 64);
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(t_acc, ra);
            tmem_ld_32x32(t_acc + 32u, rb);
            tmem_ld_wait_dep(ra);
            tmem_ld_wait_dep(rb);
            if (i > 0) {        // the projection of the previous half has read the x0 tiles (and its accumulator is complete)
                mbar_wait(&bars[1], (uint32_t)((i - 1) & 1));
                tc_fence_after();
            }
            auto piece = [&](const uint32_t (&r)[32], int c) {
                if (lane == 0) tma_store_wait_read();   // the previous store of this warp has drained the staging tile
                __syncwarp();
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const uint32_t o = (uint32_t)((part * 64 + c * 32 + ch * 8) * 4);
                    const float4 s0 = lds128f(par_addr + o), s1 = lds128f(par_addr + o + 16);
                    const float4 h0 = lds128f(par_addr + 1024 + o), h1 = lds128f(par_addr + 1024 + o + 16);
                    float v[8];
                    const int j = ch * 8;
                    v[0] = fmaf(s0.x, __uint_as_float(r[j]), h0.x);     v[1] = fmaf(s0.y, __uint_as_float(r[j + 1]), h0.y);
                    v[2] = fmaf(s0.z, __uint_as_float(r[j + 2]), h0.z); v[3] = fmaf(s0.w, __uint_as_float(r[j + 3]), h0.w);
                    v[4] = fmaf(s1.x, __uint_as_float(r[j + 4]), h1.x); v[5] = fmaf(s1.y, __uint_as_float(r[j + 5]), h1.y);
                    v[6] = fmaf(s1.z, __uint_as_float(r[j + 6]), h1.z); v[7] = fmaf(s1.w, __uint_as_float(r[j + 7]), h1.w);
                    if (p.relu) {
#pragma unroll
                        for (int u = 0; u < 8; u++) v[u] = fmaxf(v[u], 0.0f);
                    }
                    // x0 -> 16-byte chunk c*4 + ch of this row in K-block `part` of the projection's A operand
                    sts128(a2_row + ((((uint32_t)(c * 4 + ch)) ^ sw7) << 4), pack_bf16x2(v[0], v[1]) & mask, pack_bf16x2(v[2], v[3]) & mask,
                           pack_bf16x2(v[4], v[5]) & mask, pack_bf16x2(v[6], v[7]) & mask);
                    const float4 a0 = lds128f(par_addr + 2048 + o), a1 = lds128f(par_addr + 2048 + o + 16);
                    const float4 t0 = lds128f(par_addr + 3072 + o), t1 = lds128f(par_addr + 3072 + o + 16);
                    sts128(stage_addr + (uint32_t)(lane * 64) + (((uint32_t)ch ^ sw3) << 4),
                           pack_relu_bf16x2(fmaf(a0.x, v[0], t0.x), fmaf(a0.y, v[1], t0.y)) & mask,
                           pack_relu_bf16x2(fmaf(a0.z, v[2], t0.z), fmaf(a0.w, v[3], t0.w)) & mask,
                           pack_relu_bf16x2(fmaf(a1.x, v[4], t1.x), fmaf(a1.y, v[5], t1.y)) & mask,
                           pack_relu_bf16x2(fmaf(a1.z, v[6], t1.z), fmaf(a1.w, v[7], t1.w)) & mask);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d_addr(&tmOa, stage_addr, part * 64 + c * 32, row0);
                    tma_store_commit();
                }
            };
            piece(ra, 0);
            piece(rb, 1);
        }
        if (i > 0) epilogue2(i - 1);
        if (i + 1 < T) im2col(i + 1);
        fence_proxy_async();       // x0 tiles / im2col rows: generic-proxy writes that the MMAs read through the async proxy
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (i + 1 < T) issue_stem();
            // projection: D2[128 rows][128] (TMEM columns 256..383) = x0[128][256] x Wp[128][256]^T
#pragma unroll
            for (int kb = 0; kb < 4; kb++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    umma_bf16_elect<false>(tmem_base + 256u, a2_lo + (uint32_t)(kb * 1024 + k * 2), b2_lo + (uint32_t)(kb * 1024 + k * 2),
                                           idesc2, (kb | k) != 0 ? 1u : 0u);
            umma_commit_elect<false>(&bars[1]);
        }
    }
    if (T > 0) {
        mbar_wait(&bars[1], (uint32_t)((T - 1) & 1));
        tc_fence_after();
        epilogue2(T - 1);
    }
    if (lane == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- stems with up to 128 filters on tiles of whole boards (Connect4: 3x3 on 4 planes -> 128, gelu; TicTacToe-shaped
// stems with 128 filters) ------------------------------------------------------------------------------------------------
// A tile is 256 padded rows = 256 / P_pad whole boards.  K = taps x planes <= 64: the im2col tile is ONE 64-wide K-block
// (entries -1 / 0 / +1); the filters are two K-blocks, bf16 hi parts and bf16 lo parts at the same k positions, and both are
// multiplied with the same A tile (fp32 filter accuracy, 2 x ceil(K / 16) MMAs per 128-row half).  One CTA = one tile at a
// time, thread = row, two CTAs per SM so that one's epilogue (BN, activation, fp32 stream + bf16 operand stores) overlaps
// the other's im2col + MMAs.  The mma.sync stem this replaces was instruction-bound at 147 us per 4096 Connect4 leaves
// (fragment loads and scattered 8-byte stores; HBM floor 31 us).
struct StemTileArgs {
    const int32_t *count;
    int max_count;
    const int8_t *states;   // [leaf][H*W*CIN] HWC, values -1/0/+1
    int H, W, Wp, P_pad, act;   // act: 0 none, 1 relu, 2 gelu (exact-erf form)
    const uint16_t *wpack;  // [128 filters][2][64 k] bf16: hi parts | lo parts, k = tap*CIN + plane, zeros beyond taps*CIN
    const float *par;       // [5][128]: conv bias | BN scale | BN shift | scale_a | shift_a
    float *out_raw;         // activation as the blocked fp32 stream (optional)
    int has_a;              // out_a = relu(scale_a * activation + shift_a) as bf16 rows through tmOa
};

template <int KSZ, int CIN> struct TileCfg {
    static constexpr int K = KSZ * KSZ * CIN, KS = (K + 15) / 16, kh = KSZ / 2;
    static_assert(K <= 64, "one K-block");
    static constexpr int A_BYTES = 256 * 128, B_BYTES = 2 * 128 * 128, STAGE_BYTES = 8 * 2048, PAR_BYTES = 5 * 128 * 4;
    static int in_bytes(int H, int W, int P_pad) { return (256 / P_pad) * (H + 2 * kh) * (W + 2 * kh) * CIN + 16; }
    static int smem(int H, int W, int P_pad) { return 1024 + A_BYTES + B_BYTES + STAGE_BYTES + PAR_BYTES + in_bytes(H, W, P_pad) + 64; }
};

template <int KSZ, int CIN>
__global__ void __launch_bounds__(256, 2)
stem_tile_kernel(const __grid_constant__ CUtensorMap tmOa, const StemTileArgs p) {
    using Cfg = TileCfg<KSZ, CIN>;
    constexpr int KS = Cfg::KS, kh = Cfg::kh;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base, *sB = sA + Cfg::A_BYTES, *sStage = sB + Cfg::B_BYTES;
    float *s_par = (float *)(sStage + Cfg::STAGE_BYTES);
    uint64_t *bar = (uint64_t *)(s_par + 5 * 128);
    uint32_t *tmem_slot = (uint32_t *)(bar + 1);
    int8_t *s_in = (int8_t *)(bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int cnt = *p.count;
    if (cnt > p.max_count) cnt = p.max_count;
    const int bpt = 256 / p.P_pad;                       // boards per tile
    const int n_tiles = (cnt + bpt - 1) / bpt;
    const int WB = p.W + 2 * kh, cells_b = (p.H + 2 * kh) * WB;     // zero-bordered board: cells of CIN bytes
    const int ncell = p.H * p.W;

    for (int i = threadIdx.x; i < 2 * 128 * 8; i += 256) {   // filters: K-block 0 = hi parts, K-block 1 = lo parts; SWIZZLE_128B layout
        const int kb = i >> 10, r = (i >> 3) & 127, c = i & 7;
        *reinterpret_cast<uint4 *>(sB + kb * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4 *>(p.wpack + ((size_t)r * 2 + kb) * 64 + c * 8);
    }
    for (int i = threadIdx.x; i < 256 * 8; i += 256) *reinterpret_cast<uint4 *>(sA + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < 5 * 128; i += 256) s_par[i] = p.par[i];
    for (int i = threadIdx.x; i < bpt * cells_b * CIN; i += 256) s_in[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        if (p.has_a) tma_prefetch_desc(&tmOa);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int pos = threadIdx.x;                          // row of the tile owned by this thread (im2col + epilogue)
    const int gb = pos / p.P_pad, bp = pos - gb * p.P_pad;
    const int yy = bp / p.Wp - 1, xx = bp % p.Wp;
    const bool live_pos = yy >= 0 && yy < p.H && xx < p.W;
    const int q = warp & 3, sub = warp >> 2;              // TMEM lane quarter / 128-row half: rows sub*128 + q*32 + lane = pos
    const uint32_t a_row = smem_u32(sA) + (uint32_t)(pos * 128), sw7 = (uint32_t)(pos & 7), sw3 = (uint32_t)((lane >> 1) & 3);
    const uint32_t par_addr = smem_u32(s_par), stage_addr = smem_u32(sStage + warp * 2048);
    const uint32_t a_lo = umma_desc_lo(smem_u32(sA)), b_lo = umma_desc_lo(smem_u32(sB));
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    const int8_t *my_in = s_in + ((size_t)gb * cells_b + (size_t)(live_pos ? yy * WB + xx : 0)) * CIN;
    uint32_t ph = 0;

    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, ph ^= 1) {
        // ---- the tile's boards -> zero-bordered int8 tiles (boards past the batch as zeros)
        for (int i = threadIdx.x; i < bpt * ncell; i += 256) {
            const int g = i / ncell, e = i - g * ncell, y = e / p.W, x = e - y * p.W;
            const int b = t * bpt + g;
            int8_t *dst = s_in + ((size_t)g * cells_b + (size_t)((y + kh) * WB + x + kh)) * CIN;
            if (CIN == 4) *reinterpret_cast<uint32_t *>(dst) = b < cnt ? *reinterpret_cast<const uint32_t *>(p.states + ((size_t)b * ncell + e) * 4) : 0u;
            else *reinterpret_cast<uint16_t *>(dst) = b < cnt ? *reinterpret_cast<const uint16_t *>(p.states + ((size_t)b * ncell + e) * 2) : (uint16_t)0;
        }
        __syncthreads();
        const bool live = live_pos && t * bpt + gb < cnt;
        // ---- im2col row: k = tap * CIN + plane, -1 / 0 / +1 -> bf16 bits 0xBF80 / 0 / 0x3F80
        {
            uint32_t w[8 * KS];                            // KS k-steps = 2 * KS chunks of 4 words
#pragma unroll
            for (int j = 0; j < 8 * KS; j++) w[j] = 0u;
            if (live) {
#pragma unroll
                for (int tp = 0; tp < KSZ * KSZ; tp++) {
                    const int8_t *src = my_in + ((tp / KSZ) * WB + (tp % KSZ)) * CIN;
                    uint32_t cell;
                    if (CIN == 4) cell = *reinterpret_cast<const uint32_t *>(src);
                    else cell = *reinterpret_cast<const uint16_t *>(src);
#pragma unroll
                    for (int ci = 0; ci < CIN; ci++) {
                        const int v = (int)(int8_t)((cell >> (8 * ci)) & 0xffu);
                        const uint32_t h = v == 0 ? 0u : (v > 0 ? 0x3F80u : 0xBF80u);
                        const int k = tp * CIN + ci;
                        w[k >> 1] |= h << (16 * (k & 1));
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 2 * KS; c++)
                sts128(a_row + ((((uint32_t)c) ^ sw7) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
        fence_proxy_async();
        __syncthreads();
        // ---- D[256 rows][128 filters] = A[256][K] x (B_hi + B_lo)[128][K]^T : 2 halves x 2 filter parts x KS k-steps
        if (warp == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int kb = 0; kb < 2; kb++)
#pragma unroll
                    for (int k = 0; k < KS; k++)
                        umma_bf16_elect<false>(tmem_base + (uint32_t)(s * 128), a_lo + (uint32_t)(s * 1024 + k * 2),
                                               b_lo + (uint32_t)(kb * 1024 + k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_elect<false>(bar);
        }
        mbar_wait(bar, ph);
        tc_fence_after();
        // ---- epilogue: this thread's row, 4 pieces of 32 filters
        {
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 128);
            const int row0 = t * 256 + sub * 128 + q * 32;
            const uint32_t mask = live ? 0xffffffffu : 0u;
            float *outp = p.out_raw ? p.out_raw + ((((size_t)(row0 >> 5)) * 4) << 10) + (size_t)(lane * 8) : nullptr;
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(t_acc, ra);
            auto piece = [&](const uint32_t (&r)[32], int c32) {
                if (p.has_a) {
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const uint32_t o = (uint32_t)((c32 * 32 + ch * 8) * 4);
                    const float4 b0 = lds128f(par_addr + o), b1 = lds128f(par_addr + o + 16);
                    const float4 s0 = lds128f(par_addr + 512 + o), s1 = lds128f(par_addr + 512 + o + 16);
                    const float4 h0 = lds128f(par_addr + 1024 + o), h1 = lds128f(par_addr + 1024 + o + 16);
                    const int j = ch * 8;
                    float v[8];
                    v[0] = fmaf(s0.x, __uint_as_float(r[j]) + b0.x, h0.x);     v[1] = fmaf(s0.y, __uint_as_float(r[j + 1]) + b0.y, h0.y);
                    v[2] = fmaf(s0.z, __uint_as_float(r[j + 2]) + b0.z, h0.z); v[3] = fmaf(s0.w, __uint_as_float(r[j + 3]) + b0.w, h0.w);
                    v[4] = fmaf(s1.x, __uint_as_float(r[j + 4]) + b1.x, h1.x); v[5] = fmaf(s1.y, __uint_as_float(r[j + 5]) + b1.y, h1.y);
                    v[6] = fmaf(s1.z, __uint_as_float(r[j + 6]) + b1.z, h1.z); v[7] = fmaf(s1.w, __uint_as_float(r[j + 7]) + b1.w, h1.w);
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (p.act == 1) v[u] = fmaxf(v[u], 0.0f);
                        else if (p.act == 2) v[u] = gaz_small::gelu_erf(v[u]);
                        v[u] = live ? v[u] : 0.0f;
                    }
                    if (outp) gaz_conv::stg256(outp + (size_t)c32 * 1024 + ch * 256, v);
                    if (p.has_a) {
                        const float4 a0 = lds128f(par_addr + 1536 + o), a1 = lds128f(par_addr + 1536 + o + 16);
                        const float4 t0 = lds128f(par_addr + 2048 + o), t1 = lds128f(par_addr + 2048 + o + 16);
                        sts128(stage_addr + (uint32_t)(lane * 64) + (((uint32_t)ch ^ sw3) << 4),
                               pack_relu_bf16x2(fmaf(a0.x, v[0], t0.x), fmaf(a0.y, v[1], t0.y)) & mask,
                               pack_relu_bf16x2(fmaf(a0.z, v[2], t0.z), fmaf(a0.w, v[3], t0.w)) & mask,
                               pack_relu_bf16x2(fmaf(a1.x, v[4], t1.x), fmaf(a1.y, v[5], t1.y)) & mask,
                               pack_relu_bf16x2(fmaf(a1.z, v[6], t1.z), fmaf(a1.w, v[7], t1.w)) & mask);
                    }
                }
                if (p.has_a) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d_addr(&tmOa, stage_addr, c32 * 32, row0);
                        tma_store_commit();
                    }
                }
            };
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                tmem_ld_wait_dep(ra);
                tmem_ld_32x32(t_acc + (uint32_t)(c * 64 + 32), rb);
                piece(ra, 2 * c);
                tmem_ld_wait_dep(rb);
                if (c == 0) tmem_ld_32x32(t_acc + 64u, ra);
                piece(rb, 2 * c + 1);
            }
        }
        tc_fence_before();
        __syncthreads();   // TMEM, the A tile and the input tiles are free for the next tile
    }
    if (lane == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

} // namespace gaz_stem
