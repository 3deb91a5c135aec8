// gaz_core.cuh -- data layout + warp-per-tree search primitives (games, PUCT select,
// expansion, backup, re-root/compaction) for the B200 self-play engine.
//
// One warp owns one tree.  Every function below is executed by all 32 lanes of that
// warp in lock-step (warp-uniform control flow); `Coop` hides the lane/shuffle
// plumbing.  The same source is compiled a second time by tests/ with GAZ_EMUL
// (Coop of width 1, plain g++) so the CPU test-suite can exercise the exact tree
// logic without a GPU; that build is test infrastructure and is never loaded by the
// package.
//
// Reference semantics restated here (file:line relative to /root/reference):
//   MCTS.py:172-191   _get_best_PUCT_score_index      -> puct_score / select
//   MCTS.py:193-222   _PUCT_select                    -> puct_select_step
//   MCTS.py:247-294   get_terminal_actions_fn         -> terminal_scan
//   MCTS.py:367-428   _expand_with_terminal_actions   -> expand_terminal
//   MCTS.py:434-511   _expand                         -> puct_select_step + expand_finish
//   MCTS.py:513-526   _back_propagate                 -> backup
//   MCTS.py:296-365   create_expand_root              -> root_begin / expand_finish(kind=ROOT)
//   MCTS.py:620-671   _set_root / prune_tree          -> prune_step (+ in-place compaction)
//   */*.py game plugins (see per-function comments)
// Numeric rules: SURVEY.md 8a V1-V7.  This translation unit is compiled with
// -fmad=false; all float/double expressions keep the reference's evaluation order.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(GAZ_EMUL)
#define GAZ_HD __host__ __device__ __forceinline__
#define GAZ_DEVICE_BUILD 1
#else
#define GAZ_HD inline
#endif

namespace gaz {

enum { GAME_TTT = 0, GAME_C4 = 1, GAME_GOMOKU = 2 };
enum { TERM_NONE = 0, TERM_DRAW = 1, TERM_WIN = 2 };
enum { LEAF_CHILD = 0, LEAF_ROOT = 1 };
enum {
    ST_OK = 0,
    ST_NODE_OVERFLOW = 1,
    ST_SLOT_OVERFLOW = 2,
    ST_LUT_MISS = 4,
    ST_BAD_STATE = 8,
};
enum { MAXP = 225, MAXL = 256, MAXNW = 16 };
// Child-slot storage is PAGED: one pool of PAGE_SLOTS-slot pages per engine, a page table per tree.  A node's block of L <= 256
// slots never straddles a page, so a block is contiguous in memory; `slot_base` is a VIRTUAL slot index of the tree.
enum { PAGE_SHIFT = 12, PAGE_SLOTS = 1 << PAGE_SHIFT };

// ------------------------------------------------------------------ records --

struct NodeRec {          // 32 B = one DRAM sector; stats of a node live in ITS OWN record
    int32_t parent;       // node index inside the tree, -1 for a root
    uint32_t slot_base;   // first slot of this node's child block in the tree's slot arena
    uint32_t visits;      // parent.child_visits[child_id]
    float value;          // parent.child_values[child_id]
    float prior;          // parent.child_prob_priors[child_id] (PUCT) / child_logit_priors (Gumbel)
    float raw;            // parent.child_raw_values[child_id] (Gumbel)
    uint32_t meta;        // L[0:8) n_exp[8:16) child_id[16:24) action[24:32)
    uint32_t meta2;       // player[0] term[1:3) tparent[3] has_win[4] hist_len[8:16) last3[16:25)
};

GAZ_HD int nr_L(const NodeRec &r) { return r.meta & 0xff; }
GAZ_HD int nr_nexp(const NodeRec &r) { return (r.meta >> 8) & 0xff; }
GAZ_HD int nr_child_id(const NodeRec &r) { return (r.meta >> 16) & 0xff; }
GAZ_HD int nr_action(const NodeRec &r) { return (r.meta >> 24) & 0xff; }
GAZ_HD int nr_player(const NodeRec &r) { return (r.meta2 & 1) ? 1 : -1; }
GAZ_HD int nr_term(const NodeRec &r) { return (r.meta2 >> 1) & 3; }
GAZ_HD bool nr_tparent(const NodeRec &r) { return (r.meta2 >> 3) & 1; }
GAZ_HD bool nr_has_win(const NodeRec &r) { return (r.meta2 >> 4) & 1; }
GAZ_HD int nr_hist_len(const NodeRec &r) { return (r.meta2 >> 8) & 0xff; }
GAZ_HD uint32_t nr_last3(const NodeRec &r) { return (r.meta2 >> 16) & 0x1ff; }
GAZ_HD uint32_t mk_meta(int L, int nexp, int child_id, int action) {
    return (uint32_t)L | ((uint32_t)nexp << 8) | ((uint32_t)child_id << 16) | ((uint32_t)action << 24);
}
GAZ_HD uint32_t mk_meta2(int player, int term, bool tparent, bool has_win, int hist_len, uint32_t last3) {
    return (player > 0 ? 1u : 0u) | ((uint32_t)term << 1) | ((uint32_t)tparent << 3) | ((uint32_t)has_win << 4) |
           ((uint32_t)(hist_len > 255 ? 255 : hist_len) << 8) | ((last3 & 0x1ffu) << 16);
}
GAZ_HD uint32_t push_last3(uint32_t last3, int action) { return ((last3 << 3) | (uint32_t)(action & 7)) & 0x1ffu; }

struct TreeState {       // 64 B
    int32_t root;
    uint32_t root_visits;
    int32_t n_nodes;
    int32_t n_slots;
    int32_t iter;        // simulations done in the current run
    int32_t limit;       // effective iteration limit of the current run (0 = idle)
    int32_t pending;     // 1 while a leaf request of this tree is outstanding
    int32_t evals;       // evaluator calls issued by this tree
    int32_t n_pages;     // slot pages this tree owns (page_table[tree][0 .. n_pages))
    // Gumbel run state (MCTS_Gumbel.py:562-679)
    int32_t g_phase, g_ntop, g_budget, g_cur, g_done_in_child, g_curiter, g_n, g_m;
    int32_t g_state;     // GS_*
    int32_t g_best_slot; // slot of the single survivor once the run is done
    int32_t pad[5];
};
enum { GS_IDLE = 0, GS_HALVE = 1, GS_VISIT = 2, GS_DONE = 3 };

struct GameState {       // live game mirrored on the device (board = the root position)
    uint32_t board[MAXNW];
    int32_t next_player;
    int32_t hist_len;
    uint32_t last3;
    int32_t winner;      // -2 running, -1/0/1 finished
    int32_t last_action;
    int32_t pad[3];      // [0], [1]: Dirichlet-noise stream positions of the game's trees
};

struct LeafRec {         // one outstanding evaluation request
    int32_t tree;
    int32_t parent;      // node to attach to (LEAF_CHILD) or -1 (LEAF_ROOT)
    int32_t slot;
    int32_t action;
    int32_t kind;
    int32_t player;      // current_player of the node being created
    int32_t hist_len;
    uint32_t last3;
    uint32_t board[MAXNW];
};

struct View {
    int game, H, W, C, P, ncell, NW;
    int n_trees, n_games, trees_per_game, node_cap, slot_cap, gumbel;
    NodeRec *nodes;
    uint32_t *boards;
    uint32_t *slot_val;
    uint8_t *slot_act;
    uint32_t *slot_child;  // Gumbel only: child node per slot (0xffffffff = None); PUCT reuses slot_val
    int32_t *page_table;   // [n_trees][max_pages]: physical page of each virtual page of a tree
    int32_t *free_pages;   // stack of free physical pages, *free_top entries
    int32_t *free_top;
    int max_pages, pool_pages;
    TreeState *trees;
    GameState *games;
    int32_t *remap;
    int32_t *leaf_count;
    LeafRec *leaves;
    int8_t *leaf_state;
    float *policy;
    float *value;
    int32_t *status;
    float c_init, c_base;
    const double *c_lut;
    int lut_n;
    int gm_cap;            // survivor capacity per tree (Gumbel)
    uint8_t *gm_ids;
    float *gm_g;
    float c_visit, c_scale;
    double c_visit_d, c_scale_d;
    int use_softmax;
    const double *gumbel_noise; // optional injected noise [n_trees][MAXL] or null
    float dir_alpha, dir_eps;   // Dirichlet exploration noise at every evaluated node (MCTS.py:243-245,481); eps 0 = off
    uint64_t noise_seed;
    const uint64_t *tree_keys;  // optional per-tree stream keys (global game id based); null = seed ^ tree
};

// -------------------------------------------------------------------- coop ---

#if defined(__CUDA_ARCH__)
struct Coop {
    int lane;
    __device__ Coop() : lane(threadIdx.x & 31) {}
    __device__ static constexpr int width() { return 32; }
    __device__ void sync() const { __syncwarp(); }
    __device__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ uint32_t lt_mask() const { return (1u << lane) - 1u; }
    template <class T> __device__ T bcast(T v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ int sum(int v) const {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ uint64_t sum64(uint64_t v) const {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ int imin(int v) const {
        for (int o = 16; o > 0; o >>= 1) { int w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
        return v;
    }
    __device__ uint32_t umax(uint32_t v) const {
        for (int o = 16; o > 0; o >>= 1) { uint32_t w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
        return v;
    }
    __device__ float fmax_(float v) const {
        for (int o = 16; o > 0; o >>= 1) { float w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
        return v;
    }
    __device__ float fsum_(float v) const {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ float fmin_(float v) const {
        for (int o = 16; o > 0; o >>= 1) { float w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
        return v;
    }
    __device__ double dmax_(double v) const {
        for (int o = 16; o > 0; o >>= 1) { double w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
        return v;
    }
    // first-max argmax: larger score wins, ties -> smaller index; idx < 0 = no candidate
    __device__ void argmax_first(double &s, int &idx) const {
        for (int o = 16; o > 0; o >>= 1) {
            double s2 = __shfl_xor_sync(0xffffffffu, s, o);
            int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
            bool take = (i2 >= 0) && (idx < 0 || s2 > s || (s2 == s && i2 < idx));
            if (take) { s = s2; idx = i2; }
        }
    }
};
#else
struct Coop {
    int lane = 0;
    static constexpr int width() { return 1; }
    void sync() const {}
    uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    uint32_t lt_mask() const { return 0u; }
    template <class T> T bcast(T v, int) const { return v; }
    int sum(int v) const { return v; }
    uint64_t sum64(uint64_t v) const { return v; }
    int imin(int v) const { return v; }
    uint32_t umax(uint32_t v) const { return v; }
    float fmax_(float v) const { return v; }
    float fsum_(float v) const { return v; }
    float fmin_(float v) const { return v; }
    double dmax_(double v) const { return v; }
    void argmax_first(double &, int &) const {}
};
#endif

GAZ_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
GAZ_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// per-warp scratch (shared memory on the device, stack in the emulation)
struct Scratch {
    float f[MAXL];
    float f2[MAXL];
    float f3[MAXL];
    float f4[MAXL];
    uint32_t u[MAXL];
    uint64_t key[MAXL];
    uint8_t act[MAXL];
    uint8_t act2[MAXL];
    uint32_t w[2 * MAXNW];
};


// ------------------------------------------------------------------- noise ---
// Counter-based RNG (Philox-4x32-10) keyed by (seed, tree) with counter (evaluation serial, child index, draw):
// the noise a node receives depends only on which game/tree it belongs to and on how many evaluations that tree
// has made, never on the GPU count or on scheduling.  Production mode only - parity runs keep dir_eps = 0.
GAZ_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
GAZ_HD void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
GAZ_HD float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); } // (0, 1)

// Gamma(alpha, 1) by Marsaglia-Tsang (alpha < 1 boosted with U^(1/alpha)); np.random.dirichlet normalises such draws.
GAZ_HD float gamma_sample(float alpha, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1) {
    const float a = alpha < 1.0f ? alpha + 1.0f : alpha;
    const float d = a - 1.0f / 3.0f, c = 1.0f / sqrtf(9.0f * d);
    float g = d;
    uint32_t r[4];
    for (uint32_t it = 0; it < 16; it++) {
        philox4x32(c0, c1, it, 0x47414d4du, k0, k1, r);
        const float u1 = u01(r[0]), u2 = u01(r[1]), u3 = u01(r[2]);
        const float x = sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2); // Box-Muller
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        if (logf(u3) < 0.5f * x * x + d - d * v + d * logf(v)) { g = d * v; break; }
    }
    if (alpha < 1.0f) {
        philox4x32(c0, c1, 0xffffu, 0x424f4f53u, k0, k1, r);
        g *= powf(u01(r[0]), 1.0f / alpha);
    }
    return g;
}

// ------------------------------------------------------------------- games ---
// Boards are two bitboards: words [0, NW/2) = stones of player -1, [NW/2, NW) = player +1.
//   Gomoku   NW=16: 15 rows x 15 bits, two rows per word (row y -> word y/2, shift 16*(y&1))
//   Connect4 NW=4 : 64-bit column-major bitboard, bit = col*7 + height (height 0 = bottom)
//   TicTacToe NW=2: 9 bits, bit = cell
// Actions: cell index y*W+x (Gomoku/TicTacToe), column (Connect4).

GAZ_HD int pidx(int player) { return player > 0 ? 1 : 0; }

GAZ_HD uint32_t gmk_row(const uint32_t *bb, int y) { return (bb[y >> 1] >> ((y & 1) * 16)) & 0x7fffu; }
GAZ_HD uint64_t c4_bb(const uint32_t *bb) { return (uint64_t)bb[0] | ((uint64_t)bb[1] << 32); }

GAZ_HD bool cell_empty(const View &v, const uint32_t *b, int c) {
    if (v.game == GAME_GOMOKU) {
        int y = c / 15, x = c - y * 15;
        return (((gmk_row(b, y) | gmk_row(b + 8, y)) >> x) & 1u) == 0;
    }
    return (((b[0] | b[1]) >> c) & 1u) == 0; // TTT
}
GAZ_HD int c4_height(const uint32_t *b, int col) {
    uint64_t occ = c4_bb(b) | c4_bb(b + 2);
    return popc64((occ >> (col * 7)) & 0x3full);
}
// number of legal actions / is action id `a` (cell or column) legal
GAZ_HD bool action_legal(const View &v, const uint32_t *b, int a) {
    if (v.game == GAME_C4) return c4_height(b, a) < 6;
    return cell_empty(v, b, a);
}
GAZ_HD int n_action_ids(const View &v) { return v.game == GAME_C4 ? 7 : v.ncell; }

// do_action_MCTS: Gomoku.py:162-167, Tictactoe.py:219-224, Connect4.py:309-316
GAZ_HD void board_play(const View &v, uint32_t *b, int a, int player) {
    int p = pidx(player);
    if (v.game == GAME_GOMOKU) {
        int y = a / 15, x = a - y * 15;
        b[p * 8 + (y >> 1)] |= 1u << (x + (y & 1) * 16);
    } else if (v.game == GAME_C4) {
        int h = c4_height(b, a);
        int bit = a * 7 + h;
        b[p * 2 + (bit >> 5)] |= 1u << (bit & 31);
    } else {
        b[p] |= 1u << a;
    }
}

GAZ_HD int run_through(uint32_t m, int pos) { // length of the run of ones through bit `pos` (bit set)
    int up = 0, dn = 0;
    uint32_t t = m >> pos;
    while (t & 1u) { up++; t >>= 1; }
    int q = pos - 1;
    while (q >= 0 && ((m >> q) & 1u)) { dn++; q--; }
    return up + dn;
}

// check_win_MCTS for "player places at action a on board b" (b does NOT yet contain the stone).
// returns TERM_WIN / TERM_DRAW / TERM_NONE.  Gomoku.py:192-255 (>=5 through the last move, no
// draw), Connect4.py:351-411 (4 through the new stone, full board = draw),
// Tictactoe.py:273-300 (any complete line, full board = draw).
GAZ_HD int reply_result(const View &v, const uint32_t *b, int a, int player) {
    int p = pidx(player);
    if (v.game == GAME_GOMOKU) {
        const uint32_t *own = b + p * 8;
        int y = a / 15, x = a - y * 15;
        uint32_t row = gmk_row(own, y) | (1u << x);
        if (run_through(row, x) >= 5) return TERM_WIN;
        uint32_t col = 0, d1 = 0, d2 = 0;
        for (int i = -4; i <= 4; i++) {
            int yy = y + i;
            if (yy < 0 || yy > 14) continue;
            uint32_t r = gmk_row(own, yy);
            col |= ((r >> x) & 1u) << (i + 4);
            int x1 = x + i, x2 = x - i;
            if (x1 >= 0 && x1 <= 14) d1 |= ((r >> x1) & 1u) << (i + 4);
            if (x2 >= 0 && x2 <= 14) d2 |= ((r >> x2) & 1u) << (i + 4);
        }
        if (run_through(col | 16u, 4) >= 5) return TERM_WIN;
        if (run_through(d1 | 16u, 4) >= 5) return TERM_WIN;
        if (run_through(d2 | 16u, 4) >= 5) return TERM_WIN;
        return TERM_NONE;
    }
    if (v.game == GAME_C4) {
        uint64_t own = c4_bb(b + p * 2), opp = c4_bb(b + (1 - p) * 2);
        int h = popc64(((own | opp) >> (a * 7)) & 0x3full);
        own |= 1ull << (a * 7 + h);
        uint64_t m;
        m = own & (own >> 7); if (m & (m >> 14)) return TERM_WIN;  // horizontal
        m = own & (own >> 1); if (m & (m >> 2)) return TERM_WIN;   // vertical
        m = own & (own >> 6); if (m & (m >> 12)) return TERM_WIN;  // diagonal
        m = own & (own >> 8); if (m & (m >> 16)) return TERM_WIN;  // diagonal
        return popc64(own | opp) == 42 ? TERM_DRAW : TERM_NONE;
    }
    uint32_t own = b[p] | (1u << a), all = own | b[1 - p];
    uint32_t opp = b[1 - p];
    const uint32_t lines[8] = {0x007u, 0x038u, 0x1c0u, 0x049u, 0x092u, 0x124u, 0x111u, 0x054u};
    for (int i = 0; i < 8; i++)
        if ((own & lines[i]) == lines[i] || (opp & lines[i]) == lines[i]) return TERM_WIN;
    return all == 0x1ffu ? TERM_DRAW : TERM_NONE;
}

GAZ_HD int cell_value(const View &v, const uint32_t *b, int c) { // board[y][x] in {-1,0,1}
    if (v.game == GAME_GOMOKU) {
        int y = c / 15, x = c - y * 15;
        if ((gmk_row(b, y) >> x) & 1u) return -1;
        if ((gmk_row(b + 8, y) >> x) & 1u) return 1;
        return 0;
    }
    if (v.game == GAME_C4) { // c = y*7+x with y = 0 at the TOP (numpy row), height = 5-y
        int y = c / 7, x = c - y * 7;
        int bit = x * 7 + (5 - y);
        if ((c4_bb(b) >> bit) & 1ull) return -1;
        if ((c4_bb(b + 2) >> bit) & 1ull) return 1;
        return 0;
    }
    if ((b[0] >> c) & 1u) return -1;
    if ((b[1] >> c) & 1u) return 1;
    return 0;
}

// get_input_state_MCTS -> int8 HWC (Gomoku.py:173-177, Tictactoe.py:229-235, Connect4.py:327-346)
template <class CG>
GAZ_HD void encode_state(const CG &cg, const View &v, const uint32_t *b, int current_player, int hist_len,
                         uint32_t last3, int8_t *out) {
    if (v.game != GAME_C4) {
        for (int c = cg.lane; c < v.ncell; c += cg.width()) {
            out[c * 2] = (int8_t)(-current_player);
            out[c * 2 + 1] = (int8_t)cell_value(v, b, c);
        }
        return;
    }
    // Connect4: ch3 = board, ch2 / ch1 = 1 / 2 moves undone, ch0 = current_player plane, replaced by
    // "3 moves undone" once hist_len >= 4.  Undo removes the top stone of the column.
    int undo = hist_len - 1;
    if (undo > 3) undo = 3;
    if (undo < 0) undo = 0;
    uint32_t pb[3][4];
    uint32_t cur[4] = {b[0], b[1], b[2], b[3]};
    for (int i = 0; i < 3; i++) {
        if (i < undo) {
            int col = (last3 >> (3 * i)) & 7;
            uint64_t a0 = c4_bb(cur), a1 = c4_bb(cur + 2);
            int h = popc64(((a0 | a1) >> (col * 7)) & 0x3full);
            uint64_t bit = ~(1ull << (col * 7 + h - 1));
            a0 &= bit; a1 &= bit;
            cur[0] = (uint32_t)a0; cur[1] = (uint32_t)(a0 >> 32);
            cur[2] = (uint32_t)a1; cur[3] = (uint32_t)(a1 >> 32);
        }
        for (int w = 0; w < 4; w++) pb[i][w] = cur[w];
    }
    for (int c = cg.lane; c < 42; c += cg.width()) {
        out[c * 4 + 3] = (int8_t)cell_value(v, b, c);
        out[c * 4 + 2] = undo >= 1 ? (int8_t)cell_value(v, pb[0], c) : (int8_t)0;
        out[c * 4 + 1] = undo >= 2 ? (int8_t)cell_value(v, pb[1], c) : (int8_t)0;
        out[c * 4 + 0] = undo >= 3 ? (int8_t)cell_value(v, pb[2], c) : (int8_t)current_player;
    }
}

// -------------------------------------------------------------- tree access --

GAZ_HD NodeRec *node_ptr(const View &v, int tree, int n) { return v.nodes + (size_t)tree * v.node_cap + n; }
GAZ_HD uint32_t *board_ptr(const View &v, int tree, int n) {
    return v.boards + ((size_t)tree * v.node_cap + n) * v.NW;
}
// virtual slot index of a tree -> index into the pooled slot arrays
GAZ_HD size_t slot_phys(const View &v, int tree, uint32_t vslot) {
    const int32_t page = v.page_table[(size_t)tree * v.max_pages + (vslot >> PAGE_SHIFT)];
    return ((size_t)page << PAGE_SHIFT) + (vslot & (PAGE_SLOTS - 1));
}
// per-tree views of the pooled arrays, indexed by virtual slot (sv[r.slot_base + i] as before the pool was paged)
template <class T> struct SlotView {
    const View *v;
    T *arr;
    int tree;
    GAZ_HD T &operator[](uint32_t vslot) const { return arr[slot_phys(*v, tree, vslot)]; }
};
GAZ_HD SlotView<uint32_t> slotv_ptr(const View &v, int tree) { return SlotView<uint32_t>{&v, v.slot_val, tree}; }
GAZ_HD SlotView<uint8_t> slota_ptr(const View &v, int tree) { return SlotView<uint8_t>{&v, v.slot_act, tree}; }
GAZ_HD SlotView<uint32_t> slotc_ptr(const View &v, int tree) { return SlotView<uint32_t>{&v, v.slot_child, tree}; }

GAZ_HD float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
GAZ_HD uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

// child node of `r` at slot i, or -1.  PUCT: expanded children are a prefix and their node index
// overwrites the prior in slot_val (the prior moves into the child's record); Gumbel: slot_child.
GAZ_HD int child_at(const View &v, int tree, const NodeRec &r, int i) {
    if (v.gumbel) {
        uint32_t c = slotc_ptr(v, tree)[r.slot_base + i];
        return c == 0xffffffffu ? -1 : (int)c;
    }
    if (nr_tparent(r) || i < nr_nexp(r)) return (int)slotv_ptr(v, tree)[r.slot_base + i];
    return -1;
}
GAZ_HD void set_child(const View &v, int tree, uint32_t slot_base, int i, int child) {
    if (v.gumbel) slotc_ptr(v, tree)[slot_base + i] = (uint32_t)child;
    else slotv_ptr(v, tree)[slot_base + i] = (uint32_t)child;
}

template <class CG> GAZ_HD void set_status(const CG &cg, const View &v, int bit) {
    if (cg.lane == 0) {
#if defined(__CUDA_ARCH__)
        atomicOr(v.status, bit);
#else
        *v.status |= bit;
#endif
    }
}

// _back_propagate (MCTS.py:513-526): stats of node x live in x's own record.
template <class CG>
GAZ_HD void backup(const CG &cg, const View &v, int tree, int node, float value, uint32_t visits) {
    cg.sync();
    if (cg.lane == 0) {
        TreeState &ts = v.trees[tree];
        int x = node;
        while (x != ts.root && x >= 0) {
            NodeRec *r = node_ptr(v, tree, x);
            r->value = r->value + value;
            r->visits += visits;
            value = -value;
            x = r->parent;
        }
        ts.root_visits += visits;
    }
    cg.sync();
}

// Terminal look-ahead over the replies of `mover` on board b (get_terminal_actions_fn,
// MCTS.py:247-294 / MCTS_Gumbel.py:281-318).  Fills sc.act[0..k) / sc.f[0..k) (mask 1.0 = win,
// 0.0 = draw) in the reference's order: PUCT = wins first via the canonical
// argsort(mask)[::-1] (descending scan index inside each class), Gumbel = scan order.
template <class CG>
GAZ_HD int terminal_scan(const CG &cg, const View &v, const uint32_t *b, int mover, Scratch &sc, bool &any_win) {
    int nids = n_action_ids(v);
    int k = 0, nwin = 0;
    // pass 1: scan order into act2 / f2
    for (int base = 0; base < nids; base += cg.width()) {
        int a = base + cg.lane;
        int r = TERM_NONE;
        if (a < nids && action_legal(v, b, a)) r = reply_result(v, b, a, mover);
        uint32_t m = cg.ballot(r != TERM_NONE);
        uint32_t mw = cg.ballot(r == TERM_WIN);
        if (r != TERM_NONE) {
            int pos = k + popc32(m & cg.lt_mask());
            sc.act2[pos] = (uint8_t)a;
            sc.f2[pos] = r == TERM_WIN ? 1.0f : 0.0f;
        }
        k += popc32(m);
        nwin += popc32(mw);
    }
    cg.sync();
    any_win = nwin > 0;
    if (k == 0) return 0;
    if (v.gumbel) {
        for (int i = cg.lane; i < k; i += cg.width()) { sc.act[i] = sc.act2[i]; sc.f[i] = sc.f2[i]; }
    } else {
        // wins (descending scan index) then draws (descending scan index)
        for (int i = cg.lane; i < k; i += cg.width()) {
            bool win = sc.f2[i] != 0.0f;
            int after = 0; // same-class entries with a larger scan index
            for (int j = i + 1; j < k; j++) after += ((sc.f2[j] != 0.0f) == win);
            int pos = win ? after : nwin + after;
            sc.act[pos] = sc.act2[i];
            sc.f[pos] = sc.f2[i];
        }
    }
    cg.sync();
    return k;
}

// Allocate `nn` nodes and `ns` slots in the tree (warp-uniform result; -1 on overflow).
template <class CG>
GAZ_HD bool tree_alloc(const CG &cg, const View &v, int tree, int nn, int ns, int &node0, int &slot0) {
    TreeState &ts = v.trees[tree];
    node0 = ts.n_nodes;
    slot0 = ts.n_slots;
    if ((slot0 & (PAGE_SLOTS - 1)) + ns > PAGE_SLOTS) slot0 = (slot0 + PAGE_SLOTS - 1) & ~(PAGE_SLOTS - 1);   // blocks never straddle a page
    if (node0 + nn > v.node_cap) { set_status(cg, v, ST_NODE_OVERFLOW); return false; }
    if (slot0 + ns > v.slot_cap) { set_status(cg, v, ST_SLOT_OVERFLOW); return false; }
    const int need = (slot0 + ns + PAGE_SLOTS - 1) >> PAGE_SHIFT;   // ns <= PAGE_SLOTS: at most one page more than the tree owns
    cg.sync();
    int ok = 1;
    if (cg.lane == 0) {
        if (need > ts.n_pages) {   // take a page from the engine's pool (only pops run concurrently; pages return in k_release)
#if defined(__CUDA_ARCH__)
            int idx = atomicSub(v.free_top, 1) - 1;
            if (idx < 0) { atomicAdd(v.free_top, 1); ok = 0; }
#else
            int idx = --(*v.free_top);
            if (idx < 0) { ++(*v.free_top); ok = 0; }
#endif
            if (ok) { v.page_table[(size_t)tree * v.max_pages + ts.n_pages] = v.free_pages[idx]; ts.n_pages++; }
        }
        if (ok) { ts.n_nodes = node0 + nn; ts.n_slots = slot0 + ns; }
    }
    ok = cg.bcast(ok, 0);
    cg.sync();
    if (!ok) { set_status(cg, v, ST_SLOT_OVERFLOW); return false; }
    return true;
}

// _expand_with_terminal_actions (MCTS.py:367-428 / MCTS_Gumbel.py:391-453) and the terminal branch
// of create_expand_root (MCTS.py:315-344).  parent < 0 => the terminal node IS the (new) root.
// sc.act / sc.f hold the k terminal replies.  Returns the terminal parent's node index.
template <class CG>
GAZ_HD int expand_terminal(const CG &cg, const View &v, int tree, int parent, int slot, int action,
                           const uint32_t *tp_board, int tp_player, int hist_len, uint32_t last3, int k,
                           bool any_win, Scratch &sc) {
    int n0, s0;
    if (!tree_alloc(cg, v, tree, 1 + k, k, n0, s0)) return -1;
    TreeState &ts = v.trees[tree];
    const bool is_root = parent < 0;
    auto sv = slotv_ptr(v, tree);
    auto sa = slota_ptr(v, tree);
    if (cg.lane == 0) {
        NodeRec *T = node_ptr(v, tree, n0);
        T->parent = parent;
        T->slot_base = (uint32_t)s0;
        T->visits = 0;
        T->value = 0.0f;
        T->raw = 0.0f;
        T->prior = is_root ? 0.0f : u2f(sv[node_ptr(v, tree, parent)->slot_base + slot]);
        T->meta = mk_meta(k, k, is_root ? 0 : slot, action);
        T->meta2 = mk_meta2(tp_player, TERM_NONE, true, any_win, hist_len, last3);
        if (!is_root) {
            NodeRec *p = node_ptr(v, tree, parent);
            set_child(v, tree, p->slot_base, slot, n0);
            p->meta = mk_meta(nr_L(*p), v.gumbel ? nr_nexp(*p) + 1 : slot + 1, nr_child_id(*p), nr_action(*p));
        } else {
            ts.root = n0;
            ts.root_visits = 0;
        }
    }
    for (int w = cg.lane; w < v.NW; w += cg.width()) board_ptr(v, tree, n0)[w] = tp_board[w];
    for (int i = cg.lane; i < k; i += cg.width()) {
        float mask = sc.f[i];
        int c = n0 + 1 + i;
        NodeRec *r = node_ptr(v, tree, c);
        r->parent = n0;
        r->slot_base = 0;
        r->prior = any_win ? mask / (float)k : 1.0f / (float)k;
        r->raw = mask;
        if (is_root) {            // MCTS.py:331-344: each child back-propagated once with `value`
            r->visits = 1;
            r->value = any_win ? 1.0f : 0.0f;
        } else if (v.gumbel) {    // MCTS_Gumbel.py:427: stats stay zero
            r->visits = 0;
            r->value = 0.0f;
        } else {                  // MCTS.py:398-400
            r->visits = 1;
            r->value = mask;
        }
        r->meta = mk_meta(0, 0, i, sc.act[i]);
        // winner of a terminal child = the player who makes the move = -tp_player
        r->meta2 = mk_meta2(-tp_player, mask != 0.0f ? TERM_WIN : TERM_DRAW, false, false, hist_len + 1,
                            push_last3(last3, sc.act[i]));
        if (v.gumbel) sv[s0 + i] = f2u(r->prior);
        set_child(v, tree, (uint32_t)s0, i, c);
        sa[s0 + i] = sc.act[i];
    }
    cg.sync();
    if (is_root && cg.lane == 0) ts.root_visits = (uint32_t)k;
    cg.sync();
    return n0;
}

// Emit an evaluation request (the NN / host evaluator fills policy[leaf], value[leaf]).
template <class CG>
GAZ_HD void emit_leaf(const CG &cg, const View &v, int tree, int parent, int slot, int action, int kind,
                      int player, int hist_len, uint32_t last3, const uint32_t *board) {
    int li = 0;
    if (cg.lane == 0) {
#if defined(__CUDA_ARCH__)
        li = atomicAdd(v.leaf_count, 1);
#else
        li = (*v.leaf_count)++;
#endif
    }
    li = cg.bcast(li, 0);
    LeafRec *L = v.leaves + li;
    if (cg.lane == 0) {
        L->tree = tree; L->parent = parent; L->slot = slot; L->action = action; L->kind = kind;
        L->player = player; L->hist_len = hist_len; L->last3 = last3;
        v.trees[tree].pending = 1;
        v.trees[tree].evals++;
    }
    for (int w = cg.lane; w < v.NW; w += cg.width()) L->board[w] = board[w];
    encode_state(cg, v, board, player, hist_len, last3, v.leaf_state + (size_t)li * v.ncell * v.C);
    cg.sync();
}

// Expansion of `node` at `slot` up to the evaluator boundary (MCTS.py:434-466,
// MCTS_Gumbel.py:459-489).  Returns true when the simulation completed without an evaluation.
template <class CG>
GAZ_HD bool expand_begin(const CG &cg, const View &v, int tree, int node, int slot, Scratch &sc) {
    NodeRec pr = *node_ptr(v, tree, node);
    int action = slota_ptr(v, tree)[pr.slot_base + slot];
    uint32_t *cb = sc.w;
    const uint32_t *pb = board_ptr(v, tree, node);
    cg.sync();
    for (int w = cg.lane; w < v.NW; w += cg.width()) cb[w] = pb[w];
    cg.sync();
    int parent_player = nr_player(pr);
    if (cg.lane == 0) board_play(v, cb, action, -parent_player);
    cg.sync();
    bool any_win;
    int k = terminal_scan(cg, v, cb, parent_player, sc, any_win);
    int hist_len = nr_hist_len(pr) + 1;
    uint32_t last3 = push_last3(nr_last3(pr), action);
    if (k > 0) {
        int T = expand_terminal(cg, v, tree, node, slot, action, cb, -parent_player, hist_len, last3, k, any_win, sc);
        if (T >= 0) backup(cg, v, tree, T, any_win ? -(float)k : 0.0f, (uint32_t)k);
        return true;
    }
    emit_leaf(cg, v, tree, node, slot, action, LEAF_CHILD, -parent_player, hist_len, last3, cb);
    return false;
}

// PUCT score of one child (MCTS.py:181-191)
GAZ_HD double puct_score(float prior, float value, uint32_t visits, double sq, double C) {
    double U = ((double)prior * (sq / (double)((int64_t)visits + 1))) * C;
    double Q = visits > 0 ? (double)(float)((double)value / (double)visits) : (double)value;
    return Q + U;
}

GAZ_HD double puct_C(const View &v, uint32_t N) {
    if ((int)N < v.lut_n) return v.c_lut[N];
    return (double)v.c_init + log(((double)N + (double)v.c_base + 1.0) / (double)v.c_base);
}

// One PUCT simulation up to the evaluator boundary: MCTS.run loop body (MCTS.py:560-580).
template <class CG> GAZ_HD void puct_select_step(const CG &cg, const View &v, int tree, Scratch &sc) {
    TreeState &ts = v.trees[tree];
    if (ts.limit <= 0 || ts.iter >= ts.limit || ts.pending || ts.root < 0) return;
    int node = ts.root;
    uint32_t N = ts.root_visits;
    const auto sv = slotv_ptr(v, tree);
    NodeRec r = *node_ptr(v, tree, node);
    bool forced = nr_nexp(r) < nr_L(r); // "0 in root.child_visits": MCTS.py:564-570
    for (;;) {
        if (!forced && nr_tparent(r)) { // terminal parent shortcut: MCTS.py:200-208
            int pick = 0;
            if (nr_has_win(r)) {
                int L = nr_L(r);
                int best = L;
                for (int i = cg.lane; i < L; i += cg.width()) {
                    const NodeRec *c = node_ptr(v, tree, (int)sv[r.slot_base + i]);
                    if (nr_term(*c) == TERM_WIN && i < best) best = i;
                }
                best = cg.imin(best);
                pick = best < L ? best : 0; // np.random.randint -> low
            }
            int c = (int)sv[r.slot_base + pick];
            float value = nr_term(*node_ptr(v, tree, c)) == TERM_WIN ? 1.0f : 0.0f;
            backup(cg, v, tree, c, value, 1);
            if (cg.lane == 0) ts.iter++;
            cg.sync();
            return;
        }
        int nexp = nr_nexp(r), L = nr_L(r);
        int best = nexp;
        if (!forced) {
            double sq = sqrt((double)N);
            double C = puct_C(v, N);
            if ((int)N >= v.lut_n) set_status(cg, v, ST_LUT_MISS);
            int ncand = nexp < L ? nexp + 1 : nexp;
            double bs = 0.0;
            int bi = -1;
            for (int i = cg.lane; i < ncand; i += cg.width()) {
                double s;
                if (i < nexp) {
                    const NodeRec *c = node_ptr(v, tree, (int)sv[r.slot_base + i]);
                    s = puct_score(c->prior, c->value, c->visits, sq, C);
                } else {
                    s = puct_score(u2f(sv[r.slot_base + i]), 0.0f, 0u, sq, C);
                }
                if (bi < 0 || s > bs) { bs = s; bi = i; }
            }
            cg.argmax_first(bs, bi);
            best = bi;
        }
        if (best == nexp) { // expand the next child in prior order
            bool done = expand_begin(cg, v, tree, node, nexp, sc);
            if (done && cg.lane == 0) ts.iter++;
            cg.sync();
            return;
        }
        int c = (int)sv[r.slot_base + best];
        NodeRec cr = *node_ptr(v, tree, c);
        if (nr_term(cr) != TERM_NONE) { // cannot happen below a non-terminal-parent; defensive
            set_status(cg, v, ST_BAD_STATE);
            return;
        }
        N = cr.visits;
        node = c;
        r = cr;
        forced = false;
    }
}

GAZ_HD uint32_t sortable_bits(float f) {
    uint32_t b = f2u(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Second half of _expand / create_expand_root after the evaluation (MCTS.py:468-511, :346-365;
// MCTS_Gumbel.py:490-528, :375-389): mask + renormalise (V2), sort by prior (V3), create node,
// back-propagate (-value, 1).
template <class CG> GAZ_HD void expand_finish(const CG &cg, const View &v, int leaf, Scratch &sc) {
    const LeafRec L = v.leaves[leaf];
    const int tree = L.tree;
    TreeState &ts = v.trees[tree];
    const float *pol = v.policy + (size_t)leaf * v.P;
    const float val = v.value[leaf];
    // legal actions in ascending id order + their policy entries
    int nids = n_action_ids(v);
    int n = 0;
    for (int base = 0; base < nids; base += cg.width()) {
        int a = base + cg.lane;
        bool ok = a < nids && action_legal(v, L.board, a);
        uint32_t m = cg.ballot(ok);
        if (ok) {
            int pos = n + popc32(m & cg.lt_mask());
            sc.act2[pos] = (uint8_t)a;
            sc.f2[pos] = pol[a];
        }
        n += popc32(m);
    }
    cg.sync();
    if (!v.gumbel) {
        float s = 0.0f; // sequential float32 sum, every lane redundantly (bit-exact order)
        for (int i = 0; i < n; i++) s = s + sc.f2[i];
        float gsum = 0.0f;
        const bool noisy = v.dir_eps > 0.0f;
        if (noisy) { // _apply_dirichlet (MCTS.py:243-245): (1 - eps) * p + eps * Dirichlet(alpha)
            // stream position = noisy expansions this tree has made in the CURRENT game (GameState::pad, zeroed when
            // a game is seated), so the draw does not depend on which slot / GPU the game lives on
            GameState &gs = v.games[tree / v.trees_per_game];
            const uint32_t serial = (uint32_t)gs.pad[tree % v.trees_per_game];
            cg.sync();
            if (cg.lane == 0) gs.pad[tree % v.trees_per_game] = (int32_t)(serial + 1u);
            for (int i = cg.lane; i < n; i += cg.width()) {
                const uint64_t key = v.tree_keys ? v.tree_keys[tree] ^ v.noise_seed : v.noise_seed ^ (uint64_t)(uint32_t)tree;
                float g = gamma_sample(v.dir_alpha, (uint32_t)key, (uint32_t)(key >> 32), serial, (uint32_t)i);
                sc.f3[i] = g;
                gsum += g;
            }
            gsum = cg.fsum_(gsum);
            if (!(gsum > 0.0f)) gsum = 1.0f;
        }
        for (int i = cg.lane; i < n; i += cg.width()) {
            float p = sc.f2[i] / s;
            if (noisy) p = (1.0f - v.dir_eps) * p + v.dir_eps * (sc.f3[i] / gsum);
            sc.f2[i] = p;
            sc.key[i] = ((uint64_t)sortable_bits(p) << 8) | (uint64_t)i;
        }
        cg.sync();
        // rank sort, descending key: argsort(p)[::-1] with ties in descending index order
        for (int i = cg.lane; i < n; i += cg.width()) {
            uint64_t ki = sc.key[i];
            int rank = 0;
            for (int j = 0; j < n; j++) rank += sc.key[j] > ki;
            sc.f[rank] = sc.f2[i];
            sc.act[rank] = sc.act2[i];
        }
    } else {
        for (int i = cg.lane; i < n; i += cg.width()) { sc.f[i] = sc.f2[i]; sc.act[i] = sc.act2[i]; }
    }
    cg.sync();
    int n0, s0;
    if (!tree_alloc(cg, v, tree, 1, n, n0, s0)) {
        if (cg.lane == 0) { ts.pending = 0; ts.iter++; }
        return;
    }
    auto sv = slotv_ptr(v, tree);
    auto sa = slota_ptr(v, tree);
    for (int i = cg.lane; i < n; i += cg.width()) {
        sv[s0 + i] = f2u(sc.f[i]);
        sa[s0 + i] = sc.act[i];
        if (v.gumbel) slotc_ptr(v, tree)[s0 + i] = 0xffffffffu;
    }
    for (int w = cg.lane; w < v.NW; w += cg.width()) board_ptr(v, tree, n0)[w] = L.board[w];
    if (cg.lane == 0) {
        NodeRec *c = node_ptr(v, tree, n0);
        c->parent = L.parent;
        c->slot_base = (uint32_t)s0;
        c->visits = 0;
        c->value = 0.0f;
        c->raw = 0.0f;
        c->prior = 0.0f;
        c->meta = mk_meta(n, 0, L.kind == LEAF_ROOT ? 0 : L.slot, L.action);
        c->meta2 = mk_meta2(L.player, TERM_NONE, false, false, L.hist_len, L.last3);
        if (L.kind == LEAF_ROOT) {
            ts.root = n0;
            ts.root_visits = 0;
        } else {
            NodeRec *p = node_ptr(v, tree, L.parent);
            c->prior = u2f(sv[p->slot_base + L.slot]);
            set_child(v, tree, p->slot_base, L.slot, n0);
            if (v.gumbel) c->raw = val; // node.child_raw_values[index] = child_value
            p->meta = mk_meta(nr_L(*p), v.gumbel ? nr_nexp(*p) + 1 : L.slot + 1, nr_child_id(*p), nr_action(*p));
        }
        ts.pending = 0;
    }
    cg.sync();
    if (L.kind == LEAF_CHILD) {
        backup(cg, v, tree, n0, -val, 1);
        if (cg.lane == 0 && !v.gumbel) ts.iter++;
    }
    cg.sync();
}

// create_expand_root up to the evaluator boundary (MCTS.py:296-346).  Resets the tree.
template <class CG> GAZ_HD void root_begin(const CG &cg, const View &v, int tree, Scratch &sc) {
    TreeState &ts = v.trees[tree];
    const GameState &g = v.games[tree / v.trees_per_game];
    cg.sync();
    if (cg.lane == 0) {
        ts.root = -1; ts.root_visits = 0; ts.n_nodes = 0; ts.n_slots = 0; ts.pending = 0;
    }
    cg.sync();
    uint32_t *b = sc.w;
    for (int w = cg.lane; w < v.NW; w += cg.width()) b[w] = g.board[w];
    cg.sync();
    bool any_win;
    int k = terminal_scan(cg, v, b, g.next_player, sc, any_win);
    if (k > 0) {
        expand_terminal(cg, v, tree, -1, 0, 0, b, -g.next_player, g.hist_len, g.last3, k, any_win, sc);
        return;
    }
    emit_leaf(cg, v, tree, -1, 0, 0, LEAF_ROOT, -g.next_player, g.hist_len, g.last3, b);
}

// In-place compaction of the subtree under `keep_root` (nodes keep their relative order, so
// parents still precede children and slot blocks still ascend).
template <class CG> GAZ_HD void compact_tree(const CG &cg, const View &v, int tree, int keep_root) {
    TreeState &ts = v.trees[tree];
    int32_t *remap = v.remap + (size_t)tree * v.node_cap;
    const int nn = ts.n_nodes;
    const int W = cg.width();
    int count = 0;
    // pass 1: reachability + new indices
    for (int base = 0; base < nn; base += W) {
        int n = base + cg.lane;
        int parent = -1;
        bool keep = false;
        if (n < nn) {
            parent = node_ptr(v, tree, n)->parent;
            keep = n == keep_root;
            if (!keep && n > keep_root && parent >= keep_root && parent < base) keep = remap[parent] >= 0;
        }
        uint32_t m = cg.ballot(keep);
        for (int it = 0; it < W; it++) { // parents inside the same chunk
            bool k2 = keep;
            if (!keep && n < nn && n > keep_root && parent >= base && parent < n) k2 = (m >> (parent - base)) & 1u;
            uint32_t m2 = cg.ballot(k2);
            keep = k2;
            if (m2 == m) break;
            m = m2;
        }
        if (n < nn) remap[n] = keep ? count + popc32(m & cg.lt_mask()) : -1;
        count += popc32(m);
        cg.sync();
    }
    // pass 2: move node records and boards (read chunk, sync, write)
    for (int base = 0; base < nn; base += W) {
        int n = base + cg.lane;
        NodeRec r;
        uint32_t bw[MAXNW];
        int dst = -1;
        if (n < nn) {
            dst = remap[n];
            if (dst >= 0) {
                r = *node_ptr(v, tree, n);
                const uint32_t *bp = board_ptr(v, tree, n);
                for (int w = 0; w < v.NW; w++) bw[w] = bp[w];
                r.parent = (n == keep_root) ? -1 : remap[r.parent];
            }
        }
        cg.sync();
        if (dst >= 0) {
            *node_ptr(v, tree, dst) = r;
            uint32_t *bp = board_ptr(v, tree, dst);
            for (int w = 0; w < v.NW; w++) bp[w] = bw[w];
        }
        cg.sync();
    }
    // pass 3: slide slot blocks down, rewriting child indices of the expanded entries
    auto sv = slotv_ptr(v, tree);
    auto sa = slota_ptr(v, tree);
    const auto sc_ = slotc_ptr(v, tree);
    const bool has_sc = v.gumbel != 0;
    int new_slots = 0;
    for (int n = 0; n < count; n++) {
        NodeRec *r = node_ptr(v, tree, n);
        const int L = nr_L(*r);
        const int old_base = (int)r->slot_base;
        const int nexp = nr_nexp(*r);
        const bool tpar = nr_tparent(*r);
        // the same no-straddle rule as tree_alloc; packing a sub-sequence of blocks never places one later than before, so
        // the slide stays in place (write positions <= read positions) and inside the pages the tree already owns
        if ((new_slots & (PAGE_SLOTS - 1)) + L > PAGE_SLOTS) new_slots = (new_slots + PAGE_SLOTS - 1) & ~(PAGE_SLOTS - 1);
        cg.sync();
        for (int base = 0; base < L; base += W) {
            int i = base + cg.lane;
            uint32_t val = 0, ch = 0xffffffffu;
            uint8_t act = 0;
            if (i < L) {
                val = sv[old_base + i];
                act = sa[old_base + i];
                if (has_sc) {
                    ch = sc_[old_base + i];
                    if (ch != 0xffffffffu) ch = (uint32_t)remap[ch];
                } else if (tpar || i < nexp) {
                    val = (uint32_t)remap[val];
                }
            }
            cg.sync();
            if (i < L) {
                sv[new_slots + i] = val;
                sa[new_slots + i] = act;
                if (has_sc) sc_[new_slots + i] = ch;
            }
            cg.sync();
        }
        if (cg.lane == 0) r->slot_base = (uint32_t)new_slots;
        new_slots += L;
        cg.sync();
    }
    if (cg.lane == 0) { ts.n_nodes = count; ts.n_slots = new_slots; ts.root = 0; }
    cg.sync();
}

// prune_tree / _set_root (MCTS.py:620-671).  `action` was already applied to the game state.
template <class CG>
GAZ_HD void prune_step(const CG &cg, const View &v, int tree, int action, int create_new_root, Scratch &sc) {
    TreeState &ts = v.trees[tree];
    if (cg.lane == 0) { ts.limit = 0; ts.iter = 0; }
    cg.sync();
    int found = -1;
    if (!create_new_root && ts.root >= 0) {
        NodeRec r = *node_ptr(v, tree, ts.root);
        const auto sa = slota_ptr(v, tree);
        int L = nr_L(r);
        int best = MAXL;
        for (int i = cg.lane; i < L; i += cg.width())
            if (child_at(v, tree, r, i) >= 0 && sa[r.slot_base + i] == action && i < best) best = i;
        best = cg.imin(best);
        if (best < L) found = child_at(v, tree, r, best);
    }
    if (found >= 0) {
        uint32_t nv = node_ptr(v, tree, found)->visits;
        compact_tree(cg, v, tree, found);
        if (cg.lane == 0) ts.root_visits = nv;
        cg.sync();
        return;
    }
    root_begin(cg, v, tree, sc);
}


// ================================================================== Gumbel ===
// MCTS_Gumbel.py:77-148 numerics.  Arrays live in the warp scratch:
//   logits sc.f | visits sc.u | values sc.f2 | raw sc.f3 ; outputs pi' in sc.f4.
#define GAZ_EPS32 1.1920928955078125e-07

// stablemax (MCTS_Gumbel.py:77-80): element in double, rounded once, sequential f32 sum.
template <class CG> GAZ_HD void stablemax_warp(const CG &cg, const float *x, int n, float *out) {
    for (int i = cg.lane; i < n; i += cg.width()) {
        double xi = (double)x[i];
        out[i] = (float)(x[i] >= 0.0f ? xi + 1.0 : 1.0 / (1.0 - xi + GAZ_EPS32));
    }
    cg.sync();
    float s = 0.0f;
    for (int i = 0; i < n; i++) s = s + out[i];
    cg.sync();
    for (int i = cg.lane; i < n; i += cg.width()) out[i] = out[i] / s;
    cg.sync();
}

// softmax in float64 (MCTS_Gumbel.py:83-88); x holds double inputs in key[] (reinterpreted), result
// written as float32 to out.  Sequential f64 sum.
template <class CG> GAZ_HD void softmax_warp(const CG &cg, double *x, int n, float *out) {
    double mx = -1.0e300;
    for (int i = cg.lane; i < n; i += cg.width()) mx = x[i] > mx ? x[i] : mx;
    mx = cg.dmax_(mx);
    double c = -mx;
    cg.sync();
    for (int i = cg.lane; i < n; i += cg.width()) x[i] = exp(x[i] + c);
    cg.sync();
    double s = 0.0;
    for (int i = 0; i < n; i++) s = s + x[i];
    cg.sync();
    for (int i = cg.lane; i < n; i += cg.width()) out[i] = (float)(x[i] / s);
    cg.sync();
}

// Loads the per-slot arrays of node r into the scratch.  Returns max visits; *sum_visits = sum.
template <class CG>
GAZ_HD uint32_t gumbel_load(const CG &cg, const View &v, int tree, const NodeRec &r, Scratch &sc, uint64_t *sum_visits) {
    const auto sv = slotv_ptr(v, tree);
    int L = nr_L(r);
    uint32_t nb = 0;
    uint64_t s = 0;
    for (int i = cg.lane; i < L; i += cg.width()) {
        sc.f[i] = u2f(sv[r.slot_base + i]);
        int c = child_at(v, tree, r, i);
        uint32_t vis = 0;
        float val = 0.0f, raw = 0.0f;
        if (c >= 0) {
            const NodeRec *cr = node_ptr(v, tree, c);
            vis = cr->visits; val = cr->value; raw = cr->raw;
        }
        sc.u[i] = vis; sc.f2[i] = val; sc.f3[i] = raw;
        nb = vis > nb ? vis : nb;
        s += vis;
    }
    nb = cg.umax(nb);
    s = cg.sum64(s);
    cg.sync();
    *sum_visits = s;
    return nb;
}

// compute_pi (MCTS_Gumbel.py:126-148) with q_transform / compute_v_mix / rescale_q / sigma inlined.
// final_q selects the float64 q_transform variant used by run()'s final call (:655-662).
template <class CG>
GAZ_HD void compute_pi_warp(const CG &cg, const View &v, int L, uint32_t N_b, uint64_t sv, bool use_softmax,
                            bool final_q, Scratch &sc) {
    float *logits = sc.f, *values = sc.f2, *raw = sc.f3, *pi = sc.f4;
    uint32_t *vis = sc.u;
    double *d64 = reinterpret_cast<double *>(sc.key);
    // q = q_transform(mean) ; stored over values[]
    for (int i = cg.lane; i < L; i += cg.width()) {
        float q = 0.0f;
        if (vis[i] > 0) {
            float mean = (float)((double)values[i] / (double)vis[i]);
            q = final_q ? (float)(((double)mean + 1.0) / 2.0) : (mean + 1.0f) / 2.0f;
        }
        values[i] = q;
    }
    cg.sync();
    // probs -> pi[] (temporarily)
    if (use_softmax) {
        for (int i = cg.lane; i < L; i += cg.width()) d64[i] = (double)logits[i];
        cg.sync();
        softmax_warp(cg, d64, L, pi);
    } else {
        stablemax_warp(cg, logits, L, pi);
    }
    // compute_v_mix
    double sp = 0.0, wq = 0.0;
    for (int i = 0; i < L; i++) if (vis[i] > 0) sp = sp + (double)pi[i];
    for (int i = 0; i < L; i++) if (vis[i] > 0) wq = wq + (double)(float)(pi[i] * values[i]) / sp;
    cg.sync();
    float mn = 3.0e38f, mx = -3.0e38f;
    for (int i = cg.lane; i < L; i += cg.width()) {
        float vmix = (float)(((double)raw[i] + wq * (double)sv) / (double)(sv + 1));
        float cq = vis[i] > 0 ? values[i] : vmix;
        values[i] = cq; // completed_q
        mn = cq < mn ? cq : mn;
        mx = cq > mx ? cq : mx;
    }
    mn = cg.fmin_(mn);
    mx = cg.fmax_(mx);
    cg.sync();
    float den = mx - mn;
    if (!(den > (float)GAZ_EPS32)) den = (float)GAZ_EPS32;
    float scs = ((float)v.c_visit_d + (float)N_b) * (float)v.c_scale_d; // sigma scalars are float32
    if (use_softmax) {
        for (int i = cg.lane; i < L; i += cg.width()) {
            float cq = (values[i] - mn) / den;
            d64[i] = (double)logits[i] + (double)scs * (double)cq;
        }
        cg.sync();
        softmax_warp(cg, d64, L, pi);
    } else {
        for (int i = cg.lane; i < L; i += cg.width()) {
            float cq = (values[i] - mn) / den;
            raw[i] = logits[i] + scs * cq; // raw[] no longer needed
        }
        cg.sync();
        stablemax_warp(cg, raw, L, pi);
    }
}

// deterministic_selection (MCTS_Gumbel.py:226-243)
template <class CG> GAZ_HD int gumbel_pick(const CG &cg, const View &v, int tree, const NodeRec &r, Scratch &sc) {
    uint64_t sv;
    uint32_t nb = gumbel_load(cg, v, tree, r, sc, &sv);
    int L = nr_L(r);
    compute_pi_warp(cg, v, L, nb, sv, v.use_softmax != 0, false, sc);
    double bs = 0.0;
    int bi = -1;
    for (int i = cg.lane; i < L; i += cg.width()) {
        double s = (double)sc.f4[i] - (double)sc.u[i] / (double)(1 + sv);
        if (bi < 0 || s > bs) { bs = s; bi = i; }
    }
    cg.argmax_first(bs, bi);
    cg.sync();
    return bi;
}

// sequential_halving (MCTS_Gumbel.py:212-224) + the bookkeeping of run() :603-623.
template <class CG> GAZ_HD void gumbel_halve(const CG &cg, const View &v, int tree, Scratch &sc) {
    TreeState &ts = v.trees[tree];
    NodeRec r = *node_ptr(v, tree, ts.root);
    const auto sv = slotv_ptr(v, tree);
    uint8_t *ids = v.gm_ids + (size_t)tree * v.gm_cap;
    float *g = v.gm_g + (size_t)tree * v.gm_cap;
    const int m = ts.g_m, n = ts.g_n, phase = ts.g_phase;
    int n_top = ts.g_ntop;
    int L = nr_L(r);
    // N_b = root.child_visits.max()
    uint32_t nb = 0;
    for (int i = cg.lane; i < L; i += cg.width()) {
        int c = child_at(v, tree, r, i);
        uint32_t vis = c >= 0 ? node_ptr(v, tree, c)->visits : 0u;
        nb = vis > nb ? vis : nb;
    }
    nb = cg.umax(nb);
    double halved_m = (double)m / (double)(1ll << phase);
    if (halved_m < 1.0) halved_m = 1.0;
    int keep;
    if (phase == 0) {
        keep = m;
        n_top = L;
        const double *noise = v.gumbel_noise ? v.gumbel_noise + (size_t)tree * MAXL : nullptr;
        for (int i = cg.lane; i < L; i += cg.width()) {
            float lg = u2f(sv[r.slot_base + i]);
            sc.f[i] = noise ? (float)((double)lg + noise[i]) : lg; // g
            sc.f2[i] = sc.f[i];                                       // key
            sc.act2[i] = (uint8_t)i;
        }
    } else {
        keep = (int)halved_m;
        float scs = ((float)(int64_t)v.c_visit_d + (float)(int64_t)nb) * v.c_scale; // c_visit truncated to int64 (:213)
        for (int i = cg.lane; i < n_top; i += cg.width()) {
            int slot = ids[i];
            const NodeRec *cr = node_ptr(v, tree, child_at(v, tree, r, slot));
            float mean = (float)((double)cr->value / (double)cr->visits);
            float qhat = (mean + 1.0f) / 2.0f;
            sc.f[i] = g[i];
            sc.f2[i] = g[i] + scs * qhat;
            sc.act2[i] = (uint8_t)slot;
        }
    }
    cg.sync();
    if (keep > n_top) keep = n_top;
    int budget = 1;
    if (m > 1) {
        double bd = (double)n / (log2((double)m) * halved_m);
        budget = bd >= 1.0 ? (int)bd : 1;
    }
    // stable ascending rank sort of key; survivors = last `keep`
    for (int i = cg.lane; i < n_top; i += cg.width()) {
        float ki = sc.f2[i];
        int rank = 0;
        for (int j = 0; j < n_top; j++) rank += (sc.f2[j] < ki) || (sc.f2[j] == ki && j < i);
        int pos = rank - (n_top - keep);
        sc.u[i] = (uint32_t)pos;
    }
    cg.sync();
    for (int i = cg.lane; i < n_top; i += cg.width()) {
        int pos = (int)sc.u[i];
        if (pos >= 0) { ids[pos] = sc.act2[i]; g[pos] = sc.f[i]; }
    }
    cg.sync();
    n_top = keep;
    if (n_top == 2 || n_top == 3) {
        budget = (n - ts.g_curiter) / n_top;
        if (budget < 1) budget = 1;
    }
    if (cg.lane == 0) {
        ts.g_ntop = n_top;
        ts.g_budget = budget;
        ts.g_cur = 0;
        ts.g_done_in_child = -1;
        if (n_top == 1) { ts.g_state = GS_DONE; ts.g_best_slot = ids[0]; ts.limit = 0; }
        else ts.g_state = GS_VISIT;
    }
    cg.sync();
}

// One Gumbel simulation (or uncounted root-child expansion) up to the evaluator boundary:
// MCTS_Gumbel.run :601-648, select :245-260.
template <class CG> GAZ_HD void gumbel_step(const CG &cg, const View &v, int tree, Scratch &sc) {
    TreeState &ts = v.trees[tree];
    if (ts.limit <= 0 || ts.pending || ts.root < 0) return;
    for (;;) {
        int state = ts.g_state;
        cg.sync();
        if (state == GS_DONE || state == GS_IDLE) return;
        if (state == GS_HALVE) { gumbel_halve(cg, v, tree, sc); continue; }
        const uint8_t *ids = v.gm_ids + (size_t)tree * v.gm_cap;
        int cur = ts.g_cur, done = ts.g_done_in_child;
        cg.sync();
        if (cur >= ts.g_ntop) { // phase finished
            if (cg.lane == 0) { ts.g_phase++; ts.g_state = GS_HALVE; }
            cg.sync();
            continue;
        }
        int slot = ids[cur];
        NodeRec root = *node_ptr(v, tree, ts.root);
        int child = child_at(v, tree, root, slot);
        if (done < 0) {
            if (cg.lane == 0) ts.g_done_in_child = 0;
            cg.sync();
            if (child < 0) { // uncounted expansion of the root child (:626-628)
                expand_begin(cg, v, tree, ts.root, slot, sc);
                return;
            }
            done = 0;
        }
        if (done >= ts.g_budget) {
            if (cg.lane == 0) { ts.g_cur = cur + 1; ts.g_done_in_child = -1; }
            cg.sync();
            continue;
        }
        // one counted visit below root child `slot`
        int node = child;
        NodeRec r = *node_ptr(v, tree, node);
        int cslot = -1;
        bool terminal = nr_term(r) != TERM_NONE;
        while (!terminal) {
            cslot = gumbel_pick(cg, v, tree, r, sc);
            int c = child_at(v, tree, r, cslot);
            if (c < 0) break;
            NodeRec cr = *node_ptr(v, tree, c);
            node = c;
            r = cr;
            if (nr_term(cr) != TERM_NONE) terminal = true;
        }
        if (cg.lane == 0) { ts.g_done_in_child = done + 1; ts.g_curiter++; ts.iter++; }
        cg.sync();
        if (terminal) backup(cg, v, tree, node, nr_term(r) == TERM_WIN ? 1.0f : 0.0f, 1);
        else expand_begin(cg, v, tree, node, cslot, sc);
        return;
    }
}

// Final pi' of MCTS_Gumbel.run (:653-662, always the softmax branch) into out[0..L).
template <class CG> GAZ_HD void gumbel_final_pi(const CG &cg, const View &v, int tree, Scratch &sc, float *out) {
    TreeState &ts = v.trees[tree];
    NodeRec r = *node_ptr(v, tree, ts.root);
    uint64_t sv;
    uint32_t nb = gumbel_load(cg, v, tree, r, sc, &sv);
    int L = nr_L(r);
    compute_pi_warp(cg, v, L, nb, sv, true, true, sc);
    for (int i = cg.lane; i < L; i += cg.width()) out[i] = sc.f4[i];
    cg.sync();
}

// Final pi' of every tree scattered by action id (0 where illegal) - the `prob` column of MCTS_Gumbel.run's rows
// as compute_policy_improvement lays it out.
template <class CG> GAZ_HD void gumbel_final_pi_dense(const CG &cg, const View &v, int tree, Scratch &sc, float *out) {
    TreeState &ts = v.trees[tree];
    float *o = out + (size_t)tree * v.P;
    for (int i = cg.lane; i < v.P; i += cg.width()) o[i] = 0.0f;
    cg.sync();
    if (ts.root < 0) return;
    NodeRec r = *node_ptr(v, tree, ts.root);
    uint64_t sv;
    uint32_t nb = gumbel_load(cg, v, tree, r, sc, &sv);
    int L = nr_L(r);
    compute_pi_warp(cg, v, L, nb, sv, true, true, sc);
    const auto sa = slota_ptr(v, tree);
    for (int i = cg.lane; i < L; i += cg.width()) o[sa[r.slot_base + i]] = sc.f4[i];
    cg.sync();
}

} // namespace gaz
