/*
 * gaz_b200.h -- C ABI of the B200 self-play engine (libgaz_b200.so).
 *
 * The reference (subtotechnoblade/Grok_Alpha_Zero) has no FFI: its hot path is
 * Python duck typing.  Each entry point below therefore names the
 * reference Python interface it stands in for (file:line relative to the
 * reference root); INTEGRATION.md shows the ctypes stubs a maintainer adds.
 *
 * Conventions: plain pointers + sizes, caller-owned HOST buffers unless a name
 * ends in _dev; every call returns >= 0 on success (often a count) and < 0 on
 * error, with gaz_last_error() giving the message; one engine per process per
 * GPU; no hidden host threads.  Trees are numbered  game * trees_per_game + k.
 * Actions are integers: y*W+x for Gomoku/TicTacToe, the column for Connect4.
 */
#ifndef GAZ_B200_H
#define GAZ_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gaz_engine gaz_engine;

enum { GAZ_GAME_TICTACTOE = 0, GAZ_GAME_CONNECT4 = 1, GAZ_GAME_GOMOKU = 2 };
enum { GAZ_MODE_PUCT = 0, GAZ_MODE_GUMBEL = 1 };
enum { GAZ_TERM_NONE = 0, GAZ_TERM_DRAW = 1, GAZ_TERM_WIN = 2 };

typedef struct gaz_config {
    int32_t game;            /* GAZ_GAME_*                                                  */
    int32_t mode;            /* GAZ_MODE_PUCT: MCTS.py:75 ; GAZ_MODE_GUMBEL: MCTS_Gumbel.py:151 */
    int32_t n_games;         /* concurrent games (Self_Play.py:346-363 runs one per process) */
    int32_t trees_per_game;  /* 2 = mcts1/mcts2 of Self_Play.py:39-57, 1 = a single MCTS     */
    int32_t node_cap;        /* nodes per tree                                               */
    int32_t slot_cap;        /* child slots per tree (virtual limit; storage is paged, see slot_pool) */
    int32_t device;          /* CUDA device ordinal                                          */
    int32_t lut_n;           /* entries of the host-built C(N) table (MCTS.py:181-182)       */
    float c_puct_init;       /* MCTS.py:82  */
    float c_puct_base;       /* MCTS.py:83  */
    int32_t gumbel_m;        /* MCTS_Gumbel.py:159 */
    int32_t use_softmax;     /* activation_fn == "softmax" (MCTS_Gumbel.py:162,185) */
    double c_visit;          /* MCTS_Gumbel.py:160 */
    double c_scale;          /* MCTS_Gumbel.py:161 */
    int64_t slot_pool;       /* child slots in the engine-wide page pool; 0 = n_trees * slot_cap (every tree can reach its
                              * limit at once).  The reference's trees are unbounded Python objects (MCTS.py:20-72); here a
                              * tree takes 4096-slot pages from one pool as it grows and returns them when prune_tree
                              * compacts it, so the pool can be sized from the MEAN tree occupancy, not the worst case. */
} gaz_config;

const char *gaz_last_error(void);
int gaz_abi_version(void);

/* MCTS.__init__ / MCTS_Gumbel.__init__ (MCTS.py:78-132, MCTS_Gumbel.py:154-197) without the
 * root evaluation, which is gaz_new_roots(). */
int gaz_create(const gaz_config *cfg, gaz_engine **out);
void gaz_destroy(gaz_engine *e);

/* MCTS.update_hyperparams (MCTS.py:134-168) / MCTS_Gumbel.update_hyperparams (:199-210) */
int gaz_set_puct_params(gaz_engine *e, float c_puct_init, float c_puct_base);
int gaz_set_gumbel_params(gaz_engine *e, int m, double c_visit, double c_scale, int use_softmax);

/* The live `game` object the reference tree reads (game.board / next_player /
 * action_history, MCTS.py:297-312).  board = H*W int8 cells (row-major, -1/0/1);
 * hist_tail = the last <=3 actions, most recent first. */
int gaz_set_game(gaz_engine *e, int game, const int8_t *board, int next_player, const int16_t *hist_tail,
                 int hist_len);
/* the same for every game at once (Self_Play.py:346-363 starts one process per game; here one call uploads
 * all live `game` objects).  boards = n_games * H*W int8 cells; meta = n_games * 4 int32:
 * next_player, len(action_history), last three actions packed 3 bits each (most recent in bits 0-2;
 * only Connect4's input-state encoding reads them, Connect4.py:327-346), last action (-1 = none).
 * Stream-ordered: returns after enqueueing the copy + scatter kernel. */
int gaz_set_games(gaz_engine *e, const int8_t *boards, const int32_t *meta);
int gaz_reset_games(gaz_engine *e);
/* game.do_action + game.check_win for every game (Self_Play.py:142-144); actions[g] < 0 = skip.
 * winners_out[g] = -2 running / -1,0,1 (may be NULL). */
int gaz_apply_actions(gaz_engine *e, const int16_t *actions, int32_t *winners_out);
int gaz_get_game(gaz_engine *e, int game, int8_t *board_out, int32_t *info_out /* next_player, hist_len, winner */);
/* game.get_input_state() of EVERY live game (Self_Play.py:77; layouts of Gomoku.py:173-177, Connect4.py:327-346,
 * Tictactoe.py:229-235): states_out = n_games * H*W*C int8 (HWC); info_out = n_games * 4 int32:
 * next_player, len(action_history), winner (-2 running), last action id. */
int gaz_get_states(gaz_engine *e, int8_t *states_out, int32_t *info_out);
/* Dirichlet exploration noise applied on the device to the renormalised priors of every evaluated node
 * (MCTS._apply_dirichlet, MCTS.py:243-245,357,481-482): p = (1-eps)*p + eps*Dirichlet(alpha).  Counter-based
 * Philox streams keyed by (seed, tree), so a game's noise does not depend on the number of GPUs.  eps = 0 (the
 * default, used by every parity test) switches it off. */
int gaz_set_noise(gaz_engine *e, float dirichlet_alpha, float dirichlet_epsilon, uint64_t seed);
/* optional per-tree stream keys (n_trees values, e.g. the GLOBAL game id of the game a tree belongs to) so that the
 * noise of a game is the same however games are laid out over slots and GPUs; NULL = seed ^ tree index */
int gaz_set_tree_keys(gaz_engine *e, const uint64_t *keys);
/* final pi' (MCTS_Gumbel.py:653-662) of EVERY tree scattered by action id: n_trees * P floats */
int gaz_gumbel_pi_dense(gaz_engine *e, float *pi_out);

/* create_expand_root (MCTS.py:296-365): tree_mask[t] != 0 selects trees (NULL = all).
 * Returns the number of evaluation requests produced (terminal roots need none). */
int gaz_new_roots(gaz_engine *e, const uint8_t *tree_mask);

/* Start of MCTS.run / MCTS_Gumbel.run (MCTS.py:542-558, MCTS_Gumbel.py:570-599): limits[t] is
 * the iteration_limit of tree t (<= 0 = idle this run).  Budget rules (1 legal move -> 1,
 * limit < n_legal -> 3*n_legal) are applied on the device. */
int gaz_run_begin(gaz_engine *e, const int32_t *limits);

/* One simulation per running tree up to the evaluator boundary (MCTS.py:560-577: select,
 * terminal look-ahead, terminal-parent expansion).  Returns the number of leaf requests. */
int gaz_select(gaz_engine *e);
/* the `session.run(["policy","value"], {"inputs": ...})` boundary (MCTS.py:224-235):
 *   states_out: n_leaves * H*W*C int8 (get_input_state_MCTS layout, HWC); trees_out: owning tree */
int gaz_get_leaves(gaz_engine *e, int8_t *states_out, int32_t *trees_out);
/* the `depth` the reference hands to a caching session for each outstanding request (MCTS.py:224-235 `kwargs["depth"]`):
 * len(game.action_history) for a root evaluation (MCTS.py:346), len(node.action_history) of the PARENT node for a child
 * (MCTS.py:468-472); Session_Cache.Cache_Wrapper.run stores outputs only while depth < max_cache_depth (Session_Cache.py:
 * 13-26).  depths_out: n_leaves int32, same order as gaz_get_leaves. */
int gaz_get_leaf_depths(gaz_engine *e, int32_t *depths_out);
int gaz_put_evals(gaz_engine *e, const float *policy, const float *value, int n);
/* built-in deterministic evaluator shared with oracle/hash_eval.py (parity runs on the device) */
int gaz_eval_hash(gaz_engine *e, uint64_t salt, int logits);
/* second half of _expand + _back_propagate (MCTS.py:468-526) for every outstanding request */
int gaz_expand(gaz_engine *e);
/* number of trees whose run is not finished */
int gaz_remaining(gaz_engine *e);
/* n_rounds x (select -> hash evaluator -> expand) without host round trips; returns 0 */
int gaz_rounds_hash(gaz_engine *e, int n_rounds, uint64_t salt, int logits);

/* prune_tree (MCTS.py:657-671): actions[t] < 0 = leave tree t alone.  The game state must
 * already contain the action.  Returns the number of evaluation requests (fresh roots). */
int gaz_prune(gaz_engine *e, const int16_t *actions, int create_new_root);

/* root statistics = the rows of MCTS.run (MCTS.py:591-600).  Arrays sized >= 256.
 * info_out: [0] L, [1] n_expanded, [2] root_visits, [3] gumbel best slot, [4] n_nodes,
 * [5] n_slots, [6] iter, [7] evals */
int gaz_root_stats(gaz_engine *e, int tree, int16_t *actions, uint32_t *visits, float *values, float *priors,
                   float *raws, int8_t *term, int8_t *expanded, int64_t *info_out);
/* Root statistics of EVERY tree in one call, scattered by action id like compute_policy_improvement
 * (Gomoku.py:257-262, Connect4.py:414-419): visits_out / values_out = n_trees * P (0 where illegal or
 * unexpanded; values_out may be NULL); info_out = n_trees * 4 int32: root.visits, the move MCTS.run would
 * return for tau = 0 (np.argmax over child_visits, first maximum; Gumbel: the sequential-halving survivor),
 * iterations done in the current run, evaluator calls so far. */
int gaz_root_dense(gaz_engine *e, uint32_t *visits_out, float *values_out, int32_t *info_out);
/* final pi' of MCTS_Gumbel.run (:653-662) for one tree, L floats in slot order */
int gaz_gumbel_pi(gaz_engine *e, int tree, float *pi_out);
/* injected Gumbel(0,1) noise for parity runs: n_trees * 256 doubles or NULL to clear */
int gaz_set_gumbel_noise(gaz_engine *e, const double *noise);

/* CUDA-event stopwatch on the engine's stream (bench.py; torch events cannot see this stream) */
int gaz_timer_begin(gaz_engine *e);
int gaz_timer_end(gaz_engine *e, float *ms_out);
int gaz_sync(gaz_engine *e);
int gaz_status(gaz_engine *e); /* sticky error bits: 1 node overflow, 2 slot overflow, 4 LUT miss, 8 bad state */
/* pool occupancy of every tree: out = n_trees x 2 int32 (nodes in use, child slots in use) of the per-tree pools
 * gaz_config.node_cap / slot_cap.  The reference's trees are unbounded Python objects (MCTS.py:20-72); a host driver
 * uses this after gaz_prune to give a tree that could not hold one more move's search a fresh root instead. */
int gaz_tree_sizes(gaz_engine *e, int32_t *out);
/* Session_Cache.Cache_Wrapper (Session_Cache.py:4-26: evaluator outputs cached by the raw input bytes) on the device: a
 * direct-mapped table of `entries` evaluated positions in HBM.  Every round the leaf requests are looked up first (the FULL
 * input state is compared, so a hit returns exactly what the evaluator returned for that position); only the misses reach
 * the evaluator, packed into a dense batch; their outputs are stored afterwards.  Identical outputs => the searches are
 * bit-identical with and without the cache (tests/test_eval_cache.py holds the reference goldens with it on).
 * shared_scope 0: a game's entries are private to it (the key includes the game id); 1: all games share the table, like the
 * reference's one diskcache directory per generation.  Unlike Cache_Wrapper there is no depth limit and look-ups do not
 * stop at the first miss - both only bound the disk store's size and latency there.  Call before the first search round. */
int gaz_eval_cache_enable(gaz_engine *e, int64_t entries, int shared_scope);
/* out[0] look-ups, out[1] hits (evaluator calls saved), out[2] entries; zeros when the cache is off */
int gaz_eval_cache_stats(gaz_engine *e, int64_t *out);
/* the slot page pool: out[0] pages in the pool, out[1] pages free now, out[2] slots per page, out[3] page-table entries per tree */
int gaz_pool_info(gaz_engine *e, int64_t *out);
int64_t gaz_bytes_allocated(gaz_engine *e);

/* game.augment_sample for a batch of finished trajectories (Self_Play.py:174; Gomoku.py:264-303 and Tictactoe.py:322-358:
 * the 8 dihedral copies; Connect4.py:427-445: the np.fliplr pair) as a table-driven gather on `device`:
 *   states_out[a][t][j] = states[t][perm_state[a][j]],  policies_out[a][t][j] = policies[t][perm_policy[a][j]]
 * for n_pos positions of S = H*W*C state bytes and P policy entries.  The permutation tables are derived by the host from
 * the game class's own augment_sample (grok_alpha_zero_b200/Self_Play.py:augmentation_tables), so any game plugin's
 * augmentation that is a pure re-ordering runs here unchanged.  Host buffers in and out; synchronous. */
int gaz_augment(int device, const int8_t *states, const float *policies, int64_t n_pos, int S, int P, const int32_t *perm_state,
                const int32_t *perm_policy, int n_aug, int8_t *states_out, float *policies_out);

#ifdef __cplusplus
}
#endif
#endif
