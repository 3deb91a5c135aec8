/*
 * oracle/mcts_oracle.c -- CPU restatement of the reference search path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under grok_alpha_zero_b200/ may import,
 * link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * What it restates (all file:line relative to /root/reference):
 *   games    Gomoku/Gomoku.py:112-255, Connect4/Connect4.py:269-411,
 *            TicTacToe/Tictactoe.py:184-300   (plain int8 boards, no bitboards,
 *            so it is an independent check of the CUDA bitboard code)
 *   PUCT     MCTS.py:172-671   (select / expand / terminal look-ahead / backup /
 *            run budget rules / re-rooting)
 *   Gumbel   MCTS_Gumbel.py:77-733 (stablemax/softmax, q_transform, rescale_q,
 *            sigma, v_mix, pi', sequential halving, deterministic selection)
 *
 * Parity status: PINNED.  tests/golden/*.json hold outputs of the unmodified
 * reference (imported from /root/reference, see oracle/gen_golden.py) and
 * tests/test_oracle_golden.py checks this file against them bit for bit.
 *
 * Numeric rules -- compile with
 *   -O2 -ffp-contract=off -fno-fast-math   (no FMA contraction, IEEE order).
 *
 * Parity-mode conventions shared with the reference driver (gen_golden.py):
 *   - no Dirichlet / Gumbel noise; tau = 0
 *   - np.random.randint -> low  (lowest-index terminal child, MCTS.py:208)
 *   - prior ties: stable ascending sort, then reversed (MCTS.py:357,484)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXP 225
#define TERM_NONE 2 /* is_terminal is None */

enum { G_TTT = 0, G_C4 = 1, G_GOMOKU = 2 };

typedef struct {
    int id, H, W, P, C, ncell;
} GameDef;

static void game_def(int id, GameDef *g) {
    g->id = id;
    if (id == G_TTT) { g->H = 3; g->W = 3; g->P = 9; g->C = 2; }
    else if (id == G_C4) { g->H = 6; g->W = 7; g->P = 7; g->C = 4; }
    else { g->H = 15; g->W = 15; g->P = 225; g->C = 2; }
    g->ncell = g->H * g->W;
}

/* ---------------------------------------------------------------- games -- */

/* get_legal_actions_MCTS: Gomoku.py:112-119, Tictactoe.py:184-187 (argwhere,
 * row-major -> ascending cell index), Connect4.py:269-276 (columns with <6). */
static int legal_actions(const GameDef *g, const int8_t *b, int16_t *out) {
    int n = 0;
    if (g->id == G_C4) {
        for (int x = 0; x < 7; x++) {
            int s = 0;
            for (int y = 0; y < 6; y++) s += b[y * 7 + x] != 0;
            if (s < 6) out[n++] = (int16_t)x;
        }
    } else {
        for (int c = 0; c < g->ncell; c++)
            if (b[c] == 0) out[n++] = (int16_t)c;
    }
    return n;
}

/* do_action_MCTS: Gomoku.py:162-167, Tictactoe.py:219-224, Connect4.py:309-316 */
static void do_action(const GameDef *g, int8_t *b, int a, int player) {
    if (g->id == G_C4) {
        int s = 0;
        for (int y = 0; y < 6; y++) s += b[y * 7 + a] != 0;
        b[(5 - s) * 7 + a] = (int8_t)player;
    } else {
        b[a] = (int8_t)player;
    }
}

/* check_win_MCTS.  Returns player, 0 (draw) or -2 (not terminal). */
static int check_win(const GameDef *g, const int8_t *b, int player, int last) {
    if (g->id == G_GOMOKU) { /* Gomoku.py:192-255: only through the last move, no draw */
        int cx = last % 15, cy = last / 15;
        static const int dxs[4] = {1, 0, 1, -1}, dys[4] = {0, 1, 1, 1};
        for (int d = 0; d < 4; d++) {
            int fives = 0;
            for (int i = -4; i < 5; i++) {
                int nx = cx + dxs[d] * i, ny = cy + dys[d] * i;
                if (nx < 0 || nx > 14 || ny < 0 || ny > 14) continue; /* skipped, run NOT reset */
                if (b[ny * 15 + nx] == player) { if (++fives == 5) return player; }
                else fives = 0;
            }
        }
        return -2;
    }
    if (g->id == G_C4) { /* Connect4.py:351-411 */
        int x = last, y = -1;
        for (int r = 0; r < 6; r++) if (b[r * 7 + x] == player) { y = r; break; }
        if (y < 0) return -2; /* cannot happen on legal input */
        int start_x = x - 3 < 0 ? 0 : x - 3, end_x = x + 3 > 6 ? 6 : x + 3;
        int start_y = y + 3 > 5 ? 5 : y + 3, end_y = y - 3 < 0 ? 0 : y - 3;
        int count = 0;
        for (int i = start_x; i <= end_x; i++) {
            if (b[y * 7 + i] == player) { if (++count == 4) return player; } else count = 0;
        }
        count = 0;
        for (int i = start_y; i >= end_y; i--) {
            if (b[i * 7 + x] == player) { if (++count == 4) return player; } else count = 0;
        }
        count = 0;
        {
            int lo = x - start_x < start_y - y ? x - start_x : start_y - y;
            int hi = end_x - x < y - end_y ? end_x - x : y - end_y;
            for (int i = -lo; i <= hi; i++) {
                if (b[(y - i) * 7 + x + i] == player) { if (++count == 4) return player; } else count = 0;
            }
        }
        count = 0;
        {
            int lo = x - start_x < y - end_y ? x - start_x : y - end_y;
            int hi = end_x - x < start_y - y ? end_x - x : start_y - y;
            for (int i = -lo; i <= hi; i++) {
                if (b[(y + i) * 7 + x + i] == player) { if (++count == 4) return player; } else count = 0;
            }
        }
        for (int c = 0; c < 42; c++) if (b[c] == 0) return -2;
        return 0;
    }
    /* TicTacToe: Tictactoe.py:273-300 -- ANY complete line returns current_player */
    for (int r = 0; r < 3; r++)
        if (b[r * 3] != 0 && b[r * 3] == b[r * 3 + 1] && b[r * 3 + 1] == b[r * 3 + 2]) return player;
    for (int c = 0; c < 3; c++)
        if (b[c] != 0 && b[c] == b[3 + c] && b[3 + c] == b[6 + c]) return player;
    if (b[0] != 0 && b[0] == b[4] && b[4] == b[8]) return player;
    if (b[2] != 0 && b[2] == b[4] && b[4] == b[6]) return player;
    for (int c = 0; c < 9; c++) if (b[c] == 0) return -2;
    return 0;
}

/* get_input_state_MCTS -> int8 HWC.  Gomoku.py:173-177, Tictactoe.py:229-235:
 * ch0 = -current_player (side to move), ch1 = board.  Connect4.py:327-346:
 * ch3 = board, ch2/ch1 = last 1/2 moves undone, ch0 = current_player, overwritten
 * by the board with 3 moves undone once >= 4 moves were played (index wrap). */
static void input_state(const GameDef *g, const int8_t *b, int current_player,
                        const int16_t *last3, int hist_len, int8_t *out) {
    if (g->id != G_C4) {
        for (int c = 0; c < g->ncell; c++) {
            out[c * 2] = (int8_t)(-current_player);
            out[c * 2 + 1] = b[c];
        }
        return;
    }
    int8_t prev[42];
    memcpy(prev, b, 42);
    for (int c = 0; c < 42; c++) {
        out[c * 4 + 0] = (int8_t)current_player;
        out[c * 4 + 1] = 0;
        out[c * 4 + 2] = 0;
        out[c * 4 + 3] = b[c];
    }
    int undo = hist_len - 1;
    if (undo > 3) undo = 3;
    for (int i = 1; i <= undo; i++) {
        int x = last3[i - 1];
        for (int r = 0; r < 6; r++) if (prev[r * 7 + x] != 0) { prev[r * 7 + x] = 0; break; }
        for (int c = 0; c < 42; c++) out[c * 4 + (3 - i)] = prev[c];
    }
}

/* ------------------------------------------------------------ evaluator -- */

typedef void (*orc_eval_fn)(void *ctx, const int8_t *state, int n_state, float *policy, float *value);

/* Deterministic, tie-free hash evaluator shared (bit for bit) with
 * oracle/hash_eval.py (drives the reference) and the CUDA engine's parity
 * evaluator.  Integer hash -> exactly representable floats. */
static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
    z ^= z >> 27; z *= 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return z;
}
#define GOLD 0x9E3779B97F4A7C15ULL

typedef struct { int P; int logits; uint64_t salt; } HashEvalCtx;

void orc_hash_eval(void *vctx, const int8_t *state, int n_state, float *policy, float *value) {
    HashEvalCtx *c = (HashEvalCtx *)vctx;
    uint64_t acc = 0;
    for (int i = 0; i < n_state; i++)
        acc += (uint64_t)(state[i] + 2) * mix64((uint64_t)(i + 1) * GOLD);
    uint64_t h0 = mix64(acc + c->salt * 0xD1B54A32D192ED03ULL);
    for (int i = 0; i < c->P; i++) {
        uint64_t r = mix64(h0 + (uint64_t)(i + 1) * GOLD);
        uint32_t k = (uint32_t)(((r >> 52) << 8) | (uint64_t)i) + 1u; /* 1 .. 2^20, distinct per i */
        policy[i] = c->logits ? (float)k * 0x1p-17f - 4.0f : (float)k;
    }
    uint64_t rv = mix64(h0 + 0x5851F42D4C957F2DULL);
    *value = (float)(rv >> 40) * 0x1p-23f - 1.0f;
}

/* ----------------------------------------------------------------- tree -- */

typedef struct {
    int parent, child_id, action;
    int8_t player; /* current_player: the player who moved INTO this node */
    int8_t term;   /* TERM_NONE or is_terminal in {-1,0,1} */
    int L, n_exp;
    int *child;        /* node index per slot, -1 = None */
    int16_t *act;      /* action per slot */
    float *prior;      /* PUCT: sorted probs; Gumbel: logits in legal order */
    uint32_t *visits;
    float *values;
    float *raw;        /* Gumbel child_raw_values */
    int8_t board[ORC_MAXP];
    int16_t last3[3];
    int hist_len;
    uint8_t live;
} Node;

typedef struct {
    GameDef g;
    int gumbel;
    Node *nodes;
    int n_nodes, cap;
    int root;
    int64_t root_visits;
    /* PUCT params (MCTS.py:78-89) */
    float c_init, c_base;
    /* Gumbel params (MCTS_Gumbel.py:154-164) */
    int m;
    double c_visit, c_scale;
    int use_softmax;
    orc_eval_fn eval;
    void *eval_ctx;
    /* counters */
    int64_t n_evals, n_sims;
    /* last Gumbel run result */
    int gumbel_best_slot;
} Tree;

static int node_new(Tree *t) {
    if (t->n_nodes == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->nodes = (Node *)realloc(t->nodes, sizeof(Node) * (size_t)t->cap);
    }
    Node *n = &t->nodes[t->n_nodes];
    memset(n, 0, sizeof(Node));
    n->parent = -1;
    n->term = TERM_NONE;
    n->live = 1;
    return t->n_nodes++;
}

static void node_alloc_slots(Tree *t, int idx, int L) {
    Node *n = &t->nodes[idx];
    n->L = L;
    n->n_exp = 0;
    int Lc = L > 0 ? L : 1;
    n->child = (int *)malloc(sizeof(int) * Lc);
    for (int i = 0; i < L; i++) n->child[i] = -1;
    n->act = (int16_t *)calloc(Lc, sizeof(int16_t));
    n->prior = (float *)calloc(Lc, sizeof(float));
    n->visits = (uint32_t *)calloc(Lc, sizeof(uint32_t));
    n->values = (float *)calloc(Lc, sizeof(float));
    n->raw = t->gumbel ? (float *)calloc(Lc, sizeof(float)) : NULL;
}

static void node_free_slots(Node *n) {
    free(n->child); free(n->act); free(n->prior); free(n->visits); free(n->values); free(n->raw);
    n->child = NULL; n->act = NULL; n->prior = NULL; n->visits = NULL; n->values = NULL; n->raw = NULL;
    n->L = 0; n->n_exp = 0; n->live = 0;
}

Tree *orc_tree_new(int game_id, int gumbel) {
    Tree *t = (Tree *)calloc(1, sizeof(Tree));
    game_def(game_id, &t->g);
    t->gumbel = gumbel;
    t->root = -1;
    t->c_init = 2.5f; t->c_base = 19652.0f;
    t->m = 16; t->c_visit = 50.0; t->c_scale = 0.1; t->use_softmax = 1;
    return t;
}

void orc_tree_free(Tree *t) {
    if (!t) return;
    for (int i = 0; i < t->n_nodes; i++) if (t->nodes[i].live) node_free_slots(&t->nodes[i]);
    free(t->nodes);
    free(t);
}

void orc_set_eval(Tree *t, orc_eval_fn fn, void *ctx) { t->eval = fn; t->eval_ctx = ctx; }
void orc_set_puct(Tree *t, float c_init, float c_base) { t->c_init = c_init; t->c_base = c_base; }
void orc_set_gumbel(Tree *t, int m, double c_visit, double c_scale, int use_softmax) {
    t->m = m; t->c_visit = c_visit; t->c_scale = c_scale; t->use_softmax = use_softmax;
}
int64_t orc_n_evals(const Tree *t) { return t->n_evals; }
int64_t orc_n_sims(const Tree *t) { return t->n_sims; }
int orc_n_nodes(const Tree *t) { return t->n_nodes; }

static void evaluate(Tree *t, const int8_t *board, int current_player, const int16_t *last3, int hist_len,
                     float *policy, float *value) {
    int8_t st[ORC_MAXP * 4];
    input_state(&t->g, board, current_player, last3, hist_len, st);
    t->eval(t->eval_ctx, st, t->g.ncell * t->g.C, policy, value);
    t->n_evals++;
}

/* get_terminal_actions_fn: MCTS.py:247-294 / MCTS_Gumbel.py:281-318.  Scans the
 * legal replies of `mover` on `board`; PUCT sorts wins first with the canonical
 * argsort(mask)[::-1] rule, Gumbel keeps scan order. */
static int terminal_actions(const Tree *t, const int8_t *board, int mover, int16_t *tact, float *tmask) {
    int16_t legal[ORC_MAXP];
    int nl = legal_actions(&t->g, board, legal);
    int k = 0;
    int8_t tmp[ORC_MAXP];
    for (int i = 0; i < nl; i++) {
        memcpy(tmp, board, (size_t)t->g.ncell);
        do_action(&t->g, tmp, legal[i], mover);
        int r = check_win(&t->g, tmp, mover, legal[i]);
        if (r == -2) continue;
        tact[k] = legal[i];
        tmask[k] = (r == mover) ? 1.0f : 0.0f;
        k++;
    }
    if (k > 1 && !t->gumbel) {
        /* stable ascending argsort of mask, then reversed */
        int16_t a2[ORC_MAXP]; float m2[ORC_MAXP];
        int n = 0;
        for (int i = 0; i < k; i++) if (tmask[i] == 0.0f) { a2[n] = tact[i]; m2[n] = 0.0f; n++; }
        for (int i = 0; i < k; i++) if (tmask[i] != 0.0f) { a2[n] = tact[i]; m2[n] = 1.0f; n++; }
        for (int i = 0; i < k; i++) { tact[i] = a2[k - 1 - i]; tmask[i] = m2[k - 1 - i]; }
    }
    return k;
}

static void child_hist(const Node *p, int action, Node *c) {
    c->last3[0] = (int16_t)action;
    c->last3[1] = p->last3[0];
    c->last3[2] = p->last3[1];
    c->hist_len = p->hist_len + 1;
}

/* _back_propagate: MCTS.py:513-526 / MCTS_Gumbel.py:530-546 */
static void back_propagate(Tree *t, int idx, float value, uint32_t visits) {
    while (t->nodes[idx].parent >= 0 && idx != t->root) {
        int id = t->nodes[idx].child_id;
        idx = t->nodes[idx].parent;
        Node *p = &t->nodes[idx];
        p->values[id] = p->values[id] + value;
        p->visits[id] += visits;
        value = -value;
    }
    t->root_visits += visits;
}

/* mask + (optional) renormalise: get_legal_actions_policy_MCTS
 * (Gomoku.py:121-150, Connect4.py:278-299, Tictactoe.py:189-209).
 * Sequential float32 sum, then float32 divide. */
static int legal_policy(const GameDef *g, const int8_t *board, const float *policy, int normalize,
                        int16_t *acts, float *out) {
    int n = legal_actions(g, board, acts);
    for (int i = 0; i < n; i++) out[i] = policy[acts[i]];
    if (normalize) {
        float s = 0.0f;
        for (int i = 0; i < n; i++) s = s + out[i];
        for (int i = 0; i < n; i++) out[i] = out[i] / s;
    }
    return n;
}

/* argsort(p)[::-1] with the canonical tie rule (stable ascending, reversed). */
static void sort_desc(int n, int16_t *acts, float *p) {
    /* insertion sort, ascending & stable, then reverse */
    for (int i = 1; i < n; i++) {
        float v = p[i]; int16_t a = acts[i];
        int j = i;
        while (j > 0 && v < p[j - 1]) { p[j] = p[j - 1]; acts[j] = acts[j - 1]; j--; }
        p[j] = v; acts[j] = a;
    }
    for (int i = 0; i < n / 2; i++) {
        float v = p[i]; p[i] = p[n - 1 - i]; p[n - 1 - i] = v;
        int16_t a = acts[i]; acts[i] = acts[n - 1 - i]; acts[n - 1 - i] = a;
    }
}

/* create_expand_root: MCTS.py:296-365 / MCTS_Gumbel.py:320-389 */
static void new_root_l3(Tree *t, const int8_t *board, int next_player, const int16_t *l3, int hist_len) {
    int r = node_new(t);
    t->root = r;
    t->root_visits = 0;
    Node *root = &t->nodes[r];
    memcpy(root->board, board, (size_t)t->g.ncell);
    root->player = (int8_t)(-next_player);
    root->hist_len = hist_len;
    for (int i = 0; i < 3; i++) root->last3[i] = l3[i];

    int16_t tact[ORC_MAXP]; float tmask[ORC_MAXP];
    int k = terminal_actions(t, board, next_player, tact, tmask);
    if (k > 0) {
        int any_win = 0;
        for (int i = 0; i < k; i++) if (tmask[i] == 1.0f) any_win = 1;
        float value = any_win ? 1.0f : 0.0f;
        node_alloc_slots(t, r, k);
        root = &t->nodes[r];
        for (int i = 0; i < k; i++) {
            root->act[i] = tact[i];
            root->prior[i] = any_win ? tmask[i] / (float)k : 1.0f / (float)k;
            if (t->gumbel) root->raw[i] = tmask[i];
        }
        for (int i = 0; i < k; i++) {
            int c = node_new(t);
            root = &t->nodes[r];
            Node *cn = &t->nodes[c];
            cn->parent = r; cn->child_id = i; cn->action = tact[i];
            cn->player = (int8_t)(-next_player); /* sic: MCTS.py:336 */
            cn->term = (tmask[i] == 1.0f) ? (int8_t)next_player : 0;
            memcpy(cn->board, board, (size_t)t->g.ncell);
            do_action(&t->g, cn->board, tact[i], next_player);
            child_hist(root, tact[i], cn);
            root->child[i] = c;
            root->n_exp = i + 1;
            back_propagate(t, c, value, 1);
        }
        return;
    }
    float policy[ORC_MAXP], value;
    evaluate(t, board, -next_player, root->last3, hist_len, policy, &value);
    int16_t acts[ORC_MAXP]; float pr[ORC_MAXP];
    int n = legal_policy(&t->g, board, policy, !t->gumbel, acts, pr);
    if (!t->gumbel) sort_desc(n, acts, pr);
    node_alloc_slots(t, r, n);
    root = &t->nodes[r];
    for (int i = 0; i < n; i++) { root->act[i] = acts[i]; root->prior[i] = pr[i]; }
}

/* _expand_with_terminal_actions: MCTS.py:367-428 / MCTS_Gumbel.py:391-453.
 * Returns the terminal parent; *value / *visits are what gets back-propagated. */
static int expand_with_terminal(Tree *t, int node_idx, int slot, const int8_t *tp_board, int tp_action,
                                const int16_t *tact, const float *tmask, int k, float *value, uint32_t *visits) {
    int any_win = 0;
    for (int i = 0; i < k; i++) if (tmask[i] == 1.0f) any_win = 1;
    int T = node_new(t);
    Node *node = &t->nodes[node_idx];
    Node *tp = &t->nodes[T];
    tp->parent = node_idx; tp->child_id = slot; tp->action = tp_action;
    tp->player = (int8_t)(-node->player);
    memcpy(tp->board, tp_board, (size_t)t->g.ncell);
    child_hist(node, tp_action, tp);
    node_alloc_slots(t, T, k);
    node = &t->nodes[node_idx];
    tp = &t->nodes[T];
    node->child[slot] = T;
    if (!t->gumbel) node->n_exp = slot + 1;
    for (int i = 0; i < k; i++) {
        tp->act[i] = tact[i];
        tp->prior[i] = any_win ? tmask[i] / (float)k : 1.0f / (float)k;
        if (t->gumbel) {
            tp->raw[i] = tmask[i]; /* MCTS_Gumbel.py:427; visits/values stay 0 */
        } else {
            tp->values[i] = tmask[i]; /* MCTS.py:398 */
            tp->visits[i] = 1;        /* MCTS.py:400 */
        }
    }
    for (int i = 0; i < k; i++) {
        int c = node_new(t);
        node = &t->nodes[node_idx];
        tp = &t->nodes[T];
        Node *cn = &t->nodes[c];
        cn->parent = T; cn->child_id = i; cn->action = tact[i];
        cn->player = node->player;
        cn->term = (tmask[i] == 1.0f) ? node->player : 0;
        child_hist(tp, tact[i], cn);
        tp->child[i] = c;
    }
    t->nodes[T].n_exp = k;
    *value = any_win ? -(float)k : 0.0f; /* -len(terminal_mask): MCTS.py:373-376,428 */
    *visits = (uint32_t)k;
    return T;
}

/* _expand: MCTS.py:434-511 (PUCT: slot = next unexpanded in prior order) and
 * MCTS_Gumbel.py:459-528 (Gumbel: explicit slot).  Returns the node to
 * back-propagate from. */
static int expand(Tree *t, int node_idx, int slot, float *value, uint32_t *visits) {
    Node *node = &t->nodes[node_idx];
    int action = node->act[slot];
    int8_t cb[ORC_MAXP];
    memcpy(cb, node->board, (size_t)t->g.ncell);
    do_action(&t->g, cb, action, -node->player);
    int16_t tact[ORC_MAXP]; float tmask[ORC_MAXP];
    int k = terminal_actions(t, cb, node->player, tact, tmask);
    if (k > 0) return expand_with_terminal(t, node_idx, slot, cb, action, tact, tmask, k, value, visits);

    Node tmp; /* history of the child */
    child_hist(node, action, &tmp);
    float policy[ORC_MAXP], v;
    evaluate(t, cb, -node->player, tmp.last3, tmp.hist_len, policy, &v);
    int16_t acts[ORC_MAXP]; float pr[ORC_MAXP];
    int n = legal_policy(&t->g, cb, policy, !t->gumbel, acts, pr);
    if (!t->gumbel) sort_desc(n, acts, pr);
    int c = node_new(t);
    node = &t->nodes[node_idx];
    Node *cn = &t->nodes[c];
    cn->parent = node_idx; cn->child_id = slot; cn->action = action;
    cn->player = (int8_t)(-node->player);
    memcpy(cn->board, cb, (size_t)t->g.ncell);
    cn->last3[0] = tmp.last3[0]; cn->last3[1] = tmp.last3[1]; cn->last3[2] = tmp.last3[2];
    cn->hist_len = tmp.hist_len;
    node_alloc_slots(t, c, n);
    node = &t->nodes[node_idx];
    cn = &t->nodes[c];
    for (int i = 0; i < n; i++) { cn->act[i] = acts[i]; cn->prior[i] = pr[i]; }
    node->child[slot] = c;
    if (t->gumbel) node->raw[slot] = v; /* MCTS_Gumbel.py:516 */
    else node->n_exp = slot + 1;
    *value = -v;
    *visits = 1;
    return c;
}

/* ----------------------------------------------------------------- PUCT -- */

/* _get_best_PUCT_score_index: MCTS.py:172-191 */
static int puct_best(const Tree *t, const Node *n, int64_t N) {
    double sq = sqrt((double)N);
    double C = (double)t->c_init + log(((double)N + (double)t->c_base + 1.0) / (double)t->c_base);
    int best = 0;
    double bs = 0.0;
    for (int i = 0; i < n->L; i++) {
        double U = ((double)n->prior[i] * (sq / (double)((int64_t)n->visits[i] + 1))) * C;
        double Q = n->visits[i] > 0 ? (double)(float)((double)n->values[i] / (double)n->visits[i])
                                    : (double)n->values[i];
        double s = Q + U;
        if (i == 0 || s > bs) { bs = s; best = i; }
    }
    return best;
}

/* _PUCT_select: MCTS.py:193-222 */
static int puct_select(const Tree *t) {
    int idx = t->root;
    int64_t N = t->root_visits;
    for (;;) {
        const Node *n = &t->nodes[idx];
        if (n->n_exp > 0 && t->nodes[n->child[0]].term != TERM_NONE) { /* terminal parent */
            float s = 0.0f;
            for (int i = 0; i < n->L; i++) s += n->values[i];
            if (s > 0.0f) {
                for (int i = 0; i < n->n_exp; i++)
                    if (t->nodes[n->child[i]].term != 0) return n->child[i]; /* randint -> low */
            }
            return n->child[0];
        }
        int best = puct_best(t, n, N);
        if (best == n->n_exp) return idx;
        N = n->visits[best];
        idx = n->child[best];
    }
}

/* n iterations of the MCTS.run loop body (MCTS.py:560-587) */
static void puct_iterations(Tree *t, int n) {
    Node *root;
    int fully = 0;
    for (int it = 0; it < n; it++) {
        root = &t->nodes[t->root];
        if (!fully) {
            int zero = 0;
            for (int i = 0; i < root->L; i++) if (root->visits[i] == 0) { zero = 1; break; }
            if (!zero) fully = 1;
        }
        int node = fully ? puct_select(t) : t->root;
        float value; uint32_t visits;
        if (t->nodes[node].term != TERM_NONE) {
            value = (t->nodes[node].term == 1 || t->nodes[node].term == -1) ? 1.0f : 0.0f;
            visits = 1;
        } else {
            node = expand(t, node, t->nodes[node].n_exp, &value, &visits);
        }
        back_propagate(t, node, value, visits);
        t->n_sims++;
    }
}

/* MCTS.run: MCTS.py:528-587.  Returns the effective iteration count. */
int orc_puct_run(Tree *t, int iteration_limit) {
    Node *root = &t->nodes[t->root];
    int16_t tmp[ORC_MAXP];
    int n_legal = legal_actions(&t->g, root->board, tmp);
    if (n_legal == 1) iteration_limit = 1;
    else if (iteration_limit < n_legal) iteration_limit = n_legal * 3;
    puct_iterations(t, iteration_limit);
    return iteration_limit;
}

/* A slice of a longer run for bench.py's bounded CPU sample: n more iterations of the same loop body,
 * without re-applying the budget rule of MCTS.py:543-546 (the caller holds the run's limit). */
int orc_puct_iterate(Tree *t, int n) {
    puct_iterations(t, n);
    return n;
}

/* --------------------------------------------------------------- Gumbel -- */

#define EPS32 1.1920928955078125e-07 /* np.finfo(np.float32).eps */

/* stablemax: MCTS_Gumbel.py:77-80 */
static void stablemax_f32(const float *x, int n, float *out) {
    float s = 0.0f;
    for (int i = 0; i < n; i++) {
        double xi = (double)x[i];
        out[i] = (float)(x[i] >= 0.0f ? xi + 1.0 : 1.0 / (1.0 - xi + EPS32));
    }
    for (int i = 0; i < n; i++) s = s + out[i];
    for (int i = 0; i < n; i++) out[i] = out[i] / s;
}

/* softmax: MCTS_Gumbel.py:83-88 (float64) */
static void softmax_f64(const double *x, int n, double *out) {
    double mx = x[0];
    for (int i = 1; i < n; i++) if (x[i] > mx) mx = x[i];
    double c = -mx, s = 0.0;
    for (int i = 0; i < n; i++) out[i] = exp(x[i] + c);
    for (int i = 0; i < n; i++) s = s + out[i];
    for (int i = 0; i < n; i++) out[i] = out[i] / s;
}

/* sigma with float32 scalars: MCTS_Gumbel.py:107-110 */
static inline float sigma_scale_f32(float c_visit, float N_b, float c_scale) {
    return (c_visit + N_b) * c_scale;
}

/* compute_pi: MCTS_Gumbel.py:126-148, with compute_v_mix :113-124 and
 * rescale_q :100-104 inlined.  q is the q_transform'ed mean value. */
static void compute_pi(const float *raw, const float *q, const float *logits, const uint32_t *visits, int n,
                       uint32_t N_b, double c_visit, double c_scale, int use_softmax, float *pi) {
    float probs[ORC_MAXP], cq[ORC_MAXP];
    double l64[ORC_MAXP], tmp64[ORC_MAXP];
    if (use_softmax) {
        for (int i = 0; i < n; i++) l64[i] = (double)logits[i];
        softmax_f64(l64, n, tmp64);
        for (int i = 0; i < n; i++) probs[i] = (float)tmp64[i];
    } else {
        stablemax_f32(logits, n, probs);
    }
    /* compute_v_mix */
    uint64_t sv = 0;
    double sp = 0.0, wq = 0.0;
    for (int i = 0; i < n; i++) sv += visits[i];
    for (int i = 0; i < n; i++) if (visits[i] > 0) sp = sp + (double)probs[i];
    for (int i = 0; i < n; i++) if (visits[i] > 0) wq = wq + (double)(float)(probs[i] * q[i]) / sp;
    for (int i = 0; i < n; i++) {
        float vmix = (float)(((double)raw[i] + wq * (double)sv) / (double)(sv + 1));
        cq[i] = visits[i] > 0 ? q[i] : vmix;
    }
    /* rescale_q */
    float mn = cq[0], mx = cq[0];
    for (int i = 1; i < n; i++) { if (cq[i] < mn) mn = cq[i]; if (cq[i] > mx) mx = cq[i]; }
    float den = mx - mn;
    if (!(den > (float)EPS32)) den = (float)EPS32;
    for (int i = 0; i < n; i++) cq[i] = (cq[i] - mn) / den;
    float sc = sigma_scale_f32((float)c_visit, (float)N_b, (float)c_scale);
    if (use_softmax) {
        for (int i = 0; i < n; i++) l64[i] = (double)logits[i] + (double)sc * (double)cq[i];
        softmax_f64(l64, n, tmp64);
        for (int i = 0; i < n; i++) pi[i] = (float)tmp64[i];
    } else {
        float x[ORC_MAXP];
        for (int i = 0; i < n; i++) x[i] = logits[i] + sc * cq[i];
        stablemax_f32(x, n, pi);
    }
}

/* deterministic_selection: MCTS_Gumbel.py:226-243 */
static int deterministic_selection(const Tree *t, const Node *n) {
    float q[ORC_MAXP], pi[ORC_MAXP];
    uint32_t N_b = 0;
    uint64_t sv = 0;
    for (int i = 0; i < n->L; i++) {
        if (n->visits[i] > N_b) N_b = n->visits[i];
        sv += n->visits[i];
        float mean = n->visits[i] > 0 ? (float)((double)n->values[i] / (double)n->visits[i]) : -1.0f;
        q[i] = n->visits[i] > 0 ? (mean + 1.0f) / 2.0f : 0.0f; /* q_transform :91-97 */
    }
    compute_pi(n->raw, q, n->prior, n->visits, n->L, N_b, t->c_visit, t->c_scale, t->use_softmax, pi);
    int best = 0;
    double bs = 0.0;
    for (int i = 0; i < n->L; i++) {
        double s = (double)pi[i] - (double)n->visits[i] / (double)(1 + sv);
        if (i == 0 || s > bs) { bs = s; best = i; }
    }
    return best;
}

/* select: MCTS_Gumbel.py:245-260.  Returns node; *slot is the child slot when
 * the returned node must be expanded at that slot. */
static int gumbel_select(const Tree *t, int idx, int *slot) {
    for (;;) {
        const Node *n = &t->nodes[idx];
        int c = deterministic_selection(t, n);
        *slot = c;
        if (n->child[c] < 0) return idx;
        if (t->nodes[n->child[c]].term != TERM_NONE) return n->child[c];
        idx = n->child[c];
    }
}

/* stable ascending argsort of float keys (numba argsort, tie-free inputs) */
static void argsort_asc(const float *key, int n, int *idx) {
    for (int i = 0; i < n; i++) idx[i] = i;
    for (int i = 1; i < n; i++) {
        int k = idx[i]; float v = key[k];
        int j = i;
        while (j > 0 && v < key[idx[j - 1]]) { idx[j] = idx[j - 1]; j--; }
        idx[j] = k;
    }
}

/* MCTS_Gumbel.run: MCTS_Gumbel.py:562-679 with sequential_halving :212-224.
 * gumbel_noise may be NULL (use_gumbel_noise=False) or L float64 samples.
 * Returns the played action; pi_out (L floats, legal order) gets the final pi'. */
int orc_gumbel_run(Tree *t, int iteration_limit, const double *gumbel_noise, float *pi_out) {
    Node *root = &t->nodes[t->root];
    int16_t tmpa[ORC_MAXP];
    int n_legal = legal_actions(&t->g, root->board, tmpa);
    if (t->m > n_legal) t->m = n_legal; /* permanent: MCTS_Gumbel.py:581-582 */
    int m = t->m, n = iteration_limit;
    int L = root->L;
    float g[ORC_MAXP], topmean[ORC_MAXP], qhat[ORC_MAXP], key[ORC_MAXP];
    int ids[ORC_MAXP], order[ORC_MAXP];
    int n_top = L;
    for (int i = 0; i < L; i++) {
        g[i] = gumbel_noise ? (float)((double)root->prior[i] + gumbel_noise[i]) : root->prior[i];
        ids[i] = i;
        topmean[i] = root->values[i]; /* value SUMS at phase 0 (root.child_values) */
    }
    int current_iteration = 0, phase = 0;
    while (n_legal > 1) {
        /* sequential_halving */
        uint32_t N_b = 0;
        for (int i = 0; i < root->L; i++) if (root->visits[i] > N_b) N_b = root->visits[i];
        double halved_m = (double)m / (double)(1LL << phase);
        if (halved_m < 1.0) halved_m = 1.0;
        int keep;
        if (phase == 0) {
            keep = m;
            for (int i = 0; i < n_top; i++) key[i] = g[i];
        } else {
            keep = (int)halved_m;
            float sc = sigma_scale_f32((float)(int64_t)t->c_visit, (float)(int64_t)N_b, (float)t->c_scale);
            for (int i = 0; i < n_top; i++) {
                qhat[i] = (topmean[i] + 1.0f) / 2.0f; /* q_transform(values, 1, -1, 1) */
                key[i] = g[i] + sc * qhat[i];
            }
        }
        argsort_asc(key, n_top, order);
        if (keep > n_top) keep = n_top;
        int budget = 1;
        if (m > 1) { /* m == 1 divides by log2(1) = 0 in the reference (ZeroDivisionError) */
            double bd = (double)n / (log2((double)m) * halved_m);
            budget = bd >= 1.0 ? (int)bd : 1;
        }
        float g2[ORC_MAXP]; int ids2[ORC_MAXP];
        for (int i = 0; i < keep; i++) { g2[i] = g[order[n_top - keep + i]]; ids2[i] = ids[order[n_top - keep + i]]; }
        n_top = keep;
        for (int i = 0; i < n_top; i++) { g[i] = g2[i]; ids[i] = ids2[i]; }
        if (n_top == 1) break;
        if (n_top == 2 || n_top == 3) {
            budget = (n - current_iteration) / n_top;
            if (budget < 1) budget = 1;
        }
        for (int r = 0; r < n_top; r++) {
            int slot = ids[r];
            float value; uint32_t visits;
            if (t->nodes[t->root].child[slot] < 0) {
                int nd = expand(t, t->root, slot, &value, &visits);
                back_propagate(t, nd, value, visits);
            }
            for (int b = 0; b < budget; b++) {
                int node = t->nodes[t->root].child[slot];
                int cslot = -1;
                if (t->nodes[node].term == TERM_NONE) node = gumbel_select(t, node, &cslot);
                if (t->nodes[node].term != TERM_NONE) {
                    value = (t->nodes[node].term == 1 || t->nodes[node].term == -1) ? 1.0f : 0.0f;
                    visits = 1;
                } else {
                    node = expand(t, node, cslot, &value, &visits);
                }
                back_propagate(t, node, value, visits);
                current_iteration++;
                t->n_sims++;
            }
        }
        root = &t->nodes[t->root];
        for (int i = 0; i < n_top; i++)
            topmean[i] = (float)((double)root->values[ids[i]] / (double)root->visits[ids[i]]);
        phase++;
    }
    root = &t->nodes[t->root];
    if (pi_out) {
        float q[ORC_MAXP];
        uint32_t N_b = 0;
        for (int i = 0; i < root->L; i++) {
            if (root->visits[i] > N_b) N_b = root->visits[i];
            float mean = root->visits[i] > 0 ? (float)((double)root->values[i] / (double)root->visits[i]) : -1.0f;
            q[i] = root->visits[i] > 0 ? (float)(((double)mean + 1.0) / 2.0) : 0.0f;
        }
        compute_pi(root->raw, q, root->prior, root->visits, root->L, N_b, t->c_visit, t->c_scale, 1, pi_out);
    }
    t->gumbel_best_slot = ids[0];
    return root->act[ids[0]];
}

/* ---------------------------------------------------------- root access -- */

int orc_root_L(const Tree *t) { return t->nodes[t->root].L; }
int orc_root_n_exp(const Tree *t) { return t->nodes[t->root].n_exp; }
int64_t orc_root_visits(const Tree *t) { return t->root_visits; }
int orc_root_term(const Tree *t) { return t->nodes[t->root].term; }

/* per-slot root stats; term[i] = TERM_NONE(2) for non-terminal / unexpanded */
void orc_root_stats(const Tree *t, int16_t *act, uint32_t *visits, float *values, float *prior, float *raw,
                    int8_t *term, int8_t *expanded) {
    const Node *r = &t->nodes[t->root];
    for (int i = 0; i < r->L; i++) {
        act[i] = r->act[i];
        visits[i] = r->visits[i];
        values[i] = r->values[i];
        prior[i] = r->prior[i];
        if (raw) raw[i] = r->raw ? r->raw[i] : 0.0f;
        expanded[i] = r->child[i] >= 0;
        term[i] = r->child[i] >= 0 ? t->nodes[r->child[i]].term : TERM_NONE;
    }
}

/* tau = 0 choice of MCTS.run (MCTS.py:602-613): first argmax of child_visits */
int orc_puct_best_action(const Tree *t) {
    const Node *r = &t->nodes[t->root];
    int best = 0;
    for (int i = 1; i < r->L; i++) if (r->visits[i] > r->visits[best]) best = i;
    return r->act[best];
}

/* prune_tree / _set_root: MCTS.py:620-671, MCTS_Gumbel.py:681-733.  The game
 * state after `action` is derived from the root's own board + history (equal to
 * the live game's by construction). */
void orc_prune(Tree *t, int action, int create_new_root) {
    Node *root = &t->nodes[t->root];
    int mover = -root->player;
    int found = -1;
    if (!create_new_root) {
        for (int i = 0; i < root->L; i++)
            if (root->child[i] >= 0 && root->act[i] == action) { found = i; break; }
    }
    if (found >= 0) {
        int c = root->child[found];
        int64_t nv = root->visits[found];
        /* free everything not under c */
        uint8_t *keep = (uint8_t *)calloc((size_t)t->n_nodes, 1);
        keep[c] = 1;
        for (int i = c + 1; i < t->n_nodes; i++) /* parents precede children */
            if (t->nodes[i].live && t->nodes[i].parent >= 0 && keep[t->nodes[i].parent]) keep[i] = 1;
        for (int i = 0; i < t->n_nodes; i++)
            if (t->nodes[i].live && !keep[i]) node_free_slots(&t->nodes[i]);
        free(keep);
        t->root = c;
        t->root_visits = nv;
        return;
    }
    int8_t nb[ORC_MAXP];
    memcpy(nb, root->board, (size_t)t->g.ncell);
    do_action(&t->g, nb, action, mover);
    int hl = root->hist_len + 1;
    int16_t l3[3] = {(int16_t)action, root->last3[0], root->last3[1]};
    for (int i = 0; i < t->n_nodes; i++) if (t->nodes[i].live) node_free_slots(&t->nodes[i]);
    t->n_nodes = 0;
    new_root_l3(t, nb, -mover, l3, hl);
}

/* public root creation from a live game: hist = full action history, oldest first */
void orc_new_root(Tree *t, const int8_t *board, int next_player, const int16_t *hist, int hist_len) {
    int16_t l3[3];
    for (int i = 0; i < 3; i++) l3[i] = (i < hist_len) ? hist[hist_len - 1 - i] : 0;
    for (int i = 0; i < t->n_nodes; i++) if (t->nodes[i].live) node_free_slots(&t->nodes[i]);
    t->n_nodes = 0;
    new_root_l3(t, board, next_player, l3, hist_len);
}

/* standalone game helpers for the property tests */
int orc_game_legal(int game_id, const int8_t *board, int16_t *out) {
    GameDef g; game_def(game_id, &g);
    return legal_actions(&g, board, out);
}
void orc_game_do_action(int game_id, int8_t *board, int action, int player) {
    GameDef g; game_def(game_id, &g);
    do_action(&g, board, action, player);
}
int orc_game_check_win(int game_id, const int8_t *board, int player, int last) {
    GameDef g; game_def(game_id, &g);
    return check_win(&g, board, player, last);
}
void orc_game_input_state(int game_id, const int8_t *board, int current_player, const int16_t *hist, int hist_len,
                          int8_t *out) {
    GameDef g; game_def(game_id, &g);
    int16_t l3[3];
    for (int i = 0; i < 3; i++) l3[i] = (i < hist_len) ? hist[hist_len - 1 - i] : 0;
    input_state(&g, board, current_player, l3, hist_len, out);
}
void orc_root_board(const Tree *t, int8_t *out) { memcpy(out, t->nodes[t->root].board, (size_t)t->g.ncell); }
int orc_root_player(const Tree *t) { return t->nodes[t->root].player; }
