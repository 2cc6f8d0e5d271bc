// GPU-resident lock-step PV-MCTS (reference pv_mcts.py:20-95).  G independent games each own a node
// arena; every simulation is three steps shared by all games:
//   aq_mcts_select         descend from the root with PUCT (pv_mcts.py:69-78) to a leaf, rebuilding
//                          the leaf's state by applying the actions on the path (lazy State.next)
//   (leaf evaluation)      aq_leaf_eval on the G leaf states = batched model.predict (pv_mcts.py:47)
//   aq_mcts_expand_backup  create the children in legal_actions() order with their priors
//                          (pv_mcts.py:53-56) and back the value up with alternating sign (:60-66)
// The reference has no virtual loss, no Dirichlet noise and no tree reuse, so batching ACROSS games
// leaves every game's search unchanged.
//
// Arithmetic is kept as the reference computes it under NumPy >= 2 (requirements.txt pins
// numpy~=2.0.2, NEP 50 promotion): w accumulates in float64; the PUCT score is
//   float32(-w/n computed in float64) + ((float32(1.25) * p) * float32(sqrt(t))) / float32(1 + n)
// all in float32, and np.argmax takes the FIRST maximum.
#include "aq_common.cuh"

using namespace aq;

// A node is split into the 16 bytes the PUCT scan of a parent reads for EVERY child (w, prior, n) and the 16 bytes only the chosen
// child needs (links, action): the selection kernel is bound by the bytes of the children it scans (133 children per level).
struct __align__(16) MctsHot {
    double w;           // cumulative value
    float prior;        // p
    int n;              // visit count
};
struct __align__(16) MctsCold {
    int first_child;    // -1 = not expanded
    int parent;         // -1 = root
    short n_children;
    short action;       // action that leads here from the parent
    int pad;
};
static_assert(sizeof(MctsHot) == 16 && sizeof(MctsCold) == 16, "node halves must be 16 bytes");
constexpr size_t kNodeBytes = sizeof(MctsHot) + sizeof(MctsCold);

struct __align__(16) MctsGame {
    AqState root;
    int node_count;
    int leaf;       // node selected by the last aq_mcts_select
    int leaf_kind;  // 0 = evaluate with the network, 1 = terminal loss (-1), 2 = terminal draw (0)
    int overflow;   // arena exhausted (caller error: max_nodes too small)
    int pad[4];
};
static_assert(sizeof(MctsGame) == 64, "MctsGame must be 64 bytes");

static inline size_t mcts_nodes_offset(int64_t G) { return ((size_t)G * sizeof(MctsGame) + 255) & ~(size_t)255; }

__global__ void mcts_reset_kernel(MctsGame *games, MctsHot *hot, MctsCold *cold, const AqState *__restrict__ roots, int64_t G,
                                  int64_t max_nodes) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    MctsGame gm;
    gm.root = load_state(roots + g);
    gm.node_count = 1;
    gm.leaf = 0;
    gm.leaf_kind = 0;
    gm.overflow = 0;
    gm.pad[0] = gm.pad[1] = gm.pad[2] = gm.pad[3] = 0;
    games[g] = gm;
    MctsHot rh;
    rh.w = 0.0; rh.prior = 0.f; rh.n = 0;
    MctsCold rc;
    rc.first_child = -1; rc.parent = -1; rc.n_children = 0; rc.action = -1; rc.pad = 0;
    hot[g * max_nodes] = rh;  // root node: Node(state, 0), pv_mcts.py:81
    cold[g * max_nodes] = rc;
}

// Descent of game g by its warp (pv_mcts.py:69-78).  The node arrays are read through plain pointers: in the fused kernel the same
// warp has just written them (expansion + backup), so the loads must be coherent ones.
__device__ __forceinline__ void select_game(MctsGame *games, const MctsHot *hot, const MctsCold *cold, int64_t g, int lane, float c_puct,
                                            AqState *leaf_states, int32_t *leaf_kind) {
    AqState s = games[g].root;
    int node = 0, kind = 0;
    // what the descent needs from a node: first_child, n_children (cold half) and n (hot half, known from the scan of its parent)
    int first = cold[0].first_child, nc = cold[0].n_children, nn = hot[0].n;
    while (true) {
        const int tf = terminal_flags(s);  // pv_mcts.py:35-42: terminal test comes first
        if (tf) { kind = (tf & 1) ? 1 : 2; break; }
        if (first < 0 || nc <= 0) { kind = 0; break; }  // `not self.child_nodes` (None or empty)
        // t = sum(child.n) (pv_mcts.py:70-72) = n - 1: the visit that expanded this node went to no child, every later
        // visit went to exactly one (terminal children count their visits too)
        const int t = nn - 1;
        const float sq = (float)sqrt((double)t);
        float best = -INFINITY;
        int best_c = 0x7fffffff;
        int b_n = 0;  // visit count of this lane's best child
        for (int c = lane; c < nc; c += 32) {
            const MctsHot ch = hot[first + c];
            const float q = ch.n ? (float)(-ch.w / (double)ch.n) : 0.0f;
            const float u = __fdiv_rn(__fmul_rn(__fmul_rn(c_puct, ch.prior), sq), (float)(1 + ch.n));
            const float sc = __fadd_rn(q, u);
            if (sc > best || (sc == best && c < best_c) || best_c == 0x7fffffff) {
                best = sc; best_c = c;
                b_n = ch.n;
            }
        }
        // warp arg-max, lowest index among equal maxima (np.argmax)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, d);
            const bool take = (oc != 0x7fffffff) && (best_c == 0x7fffffff || ob > best || (ob == best && oc < best_c));
            if (take) { best = ob; best_c = oc; }
        }
        const int owner = best_c & 31;  // child c is examined by lane c % 32, whose own best is then the global best
        node = first + best_c;
        nn = __shfl_sync(0xffffffffu, b_n, owner);
        const MctsCold cc = cold[node];  // one 16-byte record of the chosen child
        first = cc.first_child;
        nc = cc.n_children;
        s = state_after(s, cc.action);  // lazy State.next along the path
    }
    if (lane == 0) {
        games[g].leaf = node;
        games[g].leaf_kind = kind;
        store_state(leaf_states + g, s);
        leaf_kind[g] = kind;
    }
}

// one warp per game
__global__ void __launch_bounds__(128)
mcts_select_kernel(MctsGame *games, const MctsHot *hot_all, const MctsCold *cold_all, int64_t G, int64_t max_nodes, float c_puct,
                   AqState *leaf_states, int32_t *leaf_kind) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= G) return;
    select_game(games, hot_all + g * max_nodes, cold_all + g * max_nodes, g, lane, c_puct, leaf_states, leaf_kind);
}

// Expansion of the selected leaf of game g with the evaluated priors (pv_mcts.py:47-56) and the backup of its value (:60-66).
__device__ __forceinline__ void expand_backup_game(MctsGame *games, MctsHot *hot, MctsCold *cold, int64_t g, int lane, int64_t max_nodes,
                                                   const float *priors, const float *values, const uint32_t *mask, const uint8_t *pawn) {
    const int leaf = games[g].leaf, kind = games[g].leaf_kind;
    double value;
    if (kind == 0) {
        value = (double)values[g];  // value.item(): float32 -> python float
        // children in State.legal_actions() order: ordered pawn moves, then per slot H before V
        const uint32_t *m = mask + 8 * g;
        const float *pr = priors + (int64_t)AQ_ACTIONS * g;
        const int np = pawn[8 * g];
        const int first = games[g].node_count;
        int total = np;
        // wall actions = mask bits >= 81 (word 2 holds actions 64..95); counted first to bound-check the arena
        int wl = 0;
        if (lane >= 2 && lane < 8) {
            uint32_t w = m[lane];
            if (lane == 2) w >>= 17;
            wl = __popc(w);
        }
        wl = __reduce_add_sync(0xffffffffu, wl);
        total += wl;
        bool fits = (int64_t)first + total <= max_nodes;
        if (!fits) {
            if (lane == 0) games[g].overflow = 1;
            total = 0;
        }
        if (fits) {
            MctsHot ch;
            ch.w = 0.0; ch.n = 0;
            MctsCold cc;
            cc.first_child = -1; cc.parent = leaf; cc.n_children = 0; cc.pad = 0;
            if (lane < np) {
                const int a = pawn[8 * g + 1 + lane];
                cc.action = (short)a;
                ch.prior = pr[a];
                hot[first + lane] = ch;
                cold[first + lane] = cc;
            }
            int base = first + np;
            for (int k = 0; k < 4; ++k) {
                const int c = lane + 32 * k;  // interleaved candidate index: 2*slot + (0 = H, 1 = V)
                const int slot = c >> 1;
                const int a = (c & 1) ? (AQ_SQUARES + AQ_SLOTS + slot) : (AQ_SQUARES + slot);
                const bool on = (m[a >> 5] >> (a & 31)) & 1;
                const unsigned bal = __ballot_sync(0xffffffffu, on);
                if (on) {
                    const int at = base + __popc(bal & ((1u << lane) - 1));
                    cc.action = (short)a;
                    ch.prior = pr[a];
                    hot[at] = ch;
                    cold[at] = cc;
                }
                base += __popc(bal);
            }
        }
        if (lane == 0) {
            cold[leaf].first_child = first;
            cold[leaf].n_children = (short)total;
            games[g].node_count = first + total;
        }
    } else {
        value = kind == 1 ? -1.0 : 0.0;  // pv_mcts.py:39
    }
    __syncwarp();
    if (lane == 0) {
        int node = leaf;
        while (node >= 0) {  // pv_mcts.py:50-51, 63-65: w += value; n += 1; parent gets -value
            hot[node].w += value;
            hot[node].n += 1;
            value = -value;
            node = cold[node].parent;
        }
    }
}

__global__ void __launch_bounds__(128)
mcts_expand_backup_kernel(MctsGame *games, MctsHot *hot_all, MctsCold *cold_all, int64_t G, int64_t max_nodes,
                          const float *__restrict__ priors, const float *__restrict__ values,
                          const uint32_t *__restrict__ mask, const uint8_t *__restrict__ pawn) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= G) return;
    expand_backup_game(games, hot_all + g * max_nodes, cold_all + g * max_nodes, g, lane, max_nodes, priors, values, mask, pawn);
}

// Expansion + backup of simulation i and the descent of simulation i + 1 in one launch: a game's next descent only needs that game's
// own backup, so there is no reason to wait for the other 4,095 games and for a kernel boundary in between.
__global__ void __launch_bounds__(128)
mcts_expand_select_kernel(MctsGame *games, MctsHot *hot_all, MctsCold *cold_all, int64_t G, int64_t max_nodes,
                          const float *__restrict__ priors, const float *__restrict__ values, const uint32_t *__restrict__ mask,
                          const uint8_t *__restrict__ pawn, float c_puct, AqState *leaf_states, int32_t *leaf_kind) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= G) return;
    MctsHot *hot = hot_all + g * max_nodes;
    MctsCold *cold = cold_all + g * max_nodes;
    expand_backup_game(games, hot, cold, g, lane, max_nodes, priors, values, mask, pawn);
    __syncwarp();   // orders the warp's node writes (children by all lanes, the backup by lane 0) before the descent's reads
    select_game(games, hot, cold, g, lane, c_puct, leaf_states, leaf_kind);
}

__global__ void mcts_root_counts_kernel(const MctsGame *__restrict__ games, const MctsHot *__restrict__ hot_all,
                                        const MctsCold *__restrict__ cold_all, int64_t G, int64_t max_nodes, int32_t *__restrict__ counts,
                                        int16_t *__restrict__ actions, int16_t *__restrict__ n_out,
                                        int32_t *__restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= G) return;
    const MctsHot *hot = hot_all + g * max_nodes;
    const MctsCold *cold = cold_all + g * max_nodes;
    const int first = cold[0].first_child;
    const int nc = first < 0 ? 0 : cold[0].n_children;
    for (int c = lane; c < AQ_MAX_LEGAL; c += 32) {
        counts[g * AQ_MAX_LEGAL + c] = c < nc ? hot[first + c].n : 0;
        actions[g * AQ_MAX_LEGAL + c] = c < nc ? cold[first + c].action : (int16_t)-1;
    }
    if (lane == 0) {
        n_out[g] = (int16_t)nc;
        if (overflow && games[g].overflow) atomicExch(overflow, 1);
    }
}

// ------------------------------------------------------------------------------------------
// One ply of G lock-step self-play games after their searches (reference self_play.py:47-60): for every game
//   scores = counts ** (1 / temperature) / sum  (pv_mcts.py:88-95; temperature 0 = one-hot on the FIRST maximum, np.argmax)
//   policy[action] = score for the legal actions, 0 elsewhere, over all 209 actions  (self_play.py:51-54)
//   action = one draw from `scores` (self_play.py:57, np.random.choice(legal_actions, p=scores)): inverse CDF in
//            legal_actions() order of a counter-based uniform -- a hash of (seed, game id, ply), so a game's moves do not
//            depend on which other games are still running
//   state  = state.next(action) (self_play.py:60), with its terminal flags.
// One warp per game; child c of the root lives in lane c % 32, slot c / 32.
// ------------------------------------------------------------------------------------------
constexpr int kChildSlots = (AQ_MAX_LEGAL + 31) / 32;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__global__ void __launch_bounds__(128)
selfplay_move_kernel(const AqState *__restrict__ states, const int32_t *__restrict__ counts, const int16_t *__restrict__ actions,
                     const int16_t *__restrict__ n_children, const int64_t *__restrict__ game_id, int64_t G, double inv_t, int greedy,
                     uint64_t seed, int ply, T *__restrict__ policy, int16_t *__restrict__ chosen, AqState *__restrict__ moved,
                     uint8_t *__restrict__ term) {
    __shared__ T rows[4][AQ_ACTIONS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * 4 + w;
    if (g >= G) return;
    const int n = n_children[g];
    double x[kChildSlots];
    int act[kChildSlots];
    long long best = -1;
#pragma unroll
    for (int k = 0; k < kChildSlots; ++k) {
        const int c = lane + 32 * k;
        const int cnt = c < n ? counts[g * AQ_MAX_LEGAL + c] : 0;
        act[k] = c < n ? (int)actions[g * AQ_MAX_LEGAL + c] : -1;
        x[k] = inv_t == 1.0 ? (double)cnt : (cnt > 0 ? pow((double)cnt, inv_t) : 0.0);
        if (c < n) best = max(best, ((long long)cnt << 8) | (long long)(255 - c));   // largest count, then smallest index
    }
    if (greedy) {
#pragma unroll
        for (int d = 16; d; d >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, d));
        const int arg = 255 - (int)(best & 255);
#pragma unroll
        for (int k = 0; k < kChildSlots; ++k) x[k] = (lane + 32 * k == arg && n > 0) ? 1.0 : 0.0;
    }
    // running sums in child order: a warp scan per slot on top of the slots before it (fixed order: deterministic)
    double cum[kChildSlots], base = 0.0;
#pragma unroll
    for (int k = 0; k < kChildSlots; ++k) {
        double v = x[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += up;
        }
        cum[k] = base + v;
        base = __shfl_sync(0xffffffffu, cum[k], 31);
    }
    const double total = base;
    const uint64_t bits = mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(game_id[g] + 1)) + 0xD1B54A32D192ED03ull * (uint64_t)(ply + 1));
    const double target = (double)(bits >> 11) * (1.0 / 9007199254740992.0) * total;   // u in [0, 1) times the sum
    int pick = -1, last = -1;
#pragma unroll
    for (int k = 0; k < kChildSlots; ++k) {
        const unsigned over = __ballot_sync(0xffffffffu, x[k] > 0.0 && target < cum[k]);
        const unsigned any = __ballot_sync(0xffffffffu, x[k] > 0.0);
        if (pick < 0 && over) pick = 32 * k + __ffs(over) - 1;
        if (any) last = 32 * k + 31 - __clz(any);
    }
    if (pick < 0) pick = last;   // target == total after rounding: the last child with a positive score
    int a = -1;
#pragma unroll
    for (int k = 0; k < kChildSlots; ++k) {
        const int v = __shfl_sync(0xffffffffu, act[k], pick & 31);
        if ((pick >> 5) == k) a = v;
    }
    // dense policy row, staged so that the global stores are consecutive
    T *row = rows[w];
    for (int i = lane; i < AQ_ACTIONS; i += 32) row[i] = (T)0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kChildSlots; ++k)
        if (act[k] >= 0) row[act[k]] = (T)(total > 0.0 ? x[k] / total : 0.0);   // x / sum(xs), pv_mcts.py:94
    __syncwarp();
    for (int i = lane; i < AQ_ACTIONS; i += 32) policy[g * AQ_ACTIONS + i] = row[i];
    if (lane == 0) {
        AqState t = load_state(states + g);
        int flags = 2;   // a root without children cannot move: reported as a draw (does not happen for non-terminal roots)
        if (pick >= 0) {
            t = state_after(t, a);
            flags = terminal_flags(t);
        }
        store_state(moved + g, t);
        term[g] = (uint8_t)flags;
        if (chosen) chosen[g] = (int16_t)a;
    }
}

// Survivors keep their order (games that ended leave; self_play.py:45 `while not state.is_done()`); one CTA, a block scan per
// 1,024 games.  For a game that ended: final_flags[game] = terminal flags of its last state, final_plies[game] = its length.
__global__ void __launch_bounds__(1024)
selfplay_compact_kernel(const AqState *__restrict__ moved, const uint8_t *__restrict__ term, const int64_t *__restrict__ game_id,
                        int64_t G, int ply, AqState *__restrict__ next_states, int64_t *__restrict__ next_game_id,
                        uint8_t *__restrict__ final_flags, int64_t *__restrict__ final_plies, int32_t *__restrict__ alive_count) {
    __shared__ int warp_total[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int base = 0;
    for (int64_t start = 0; start < G; start += 1024) {
        const int64_t i = start + threadIdx.x;
        const int flags = i < G ? (int)term[i] : 1;
        const bool alive = i < G && flags == 0;
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) warp_total[w] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
        for (int v = 0; v < 32; ++v) {
            const int t = warp_total[v];
            before += v < w ? t : 0;
            all += t;
        }
        if (i < G) {
            const int64_t id = game_id[i];
            if (alive) {
                const int64_t o = base + before + __popc(m & ((1u << lane) - 1u));
                store_state(next_states + o, load_state(moved + i));
                next_game_id[o] = id;
            } else {
                final_flags[id] = (uint8_t)flags;
                final_plies[id] = ply + 1;
            }
        }
        base += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *alive_count = base;
}

// ------------------------------------------------------------------------------------------
extern "C" int64_t aq_mcts_ws_bytes(int64_t G, int64_t max_nodes) {
    return (int64_t)(mcts_nodes_offset(G) + (size_t)G * (size_t)max_nodes * kNodeBytes);
}

static inline MctsGame *games_of(void *ws) { return reinterpret_cast<MctsGame *>(ws); }
// workspace: games | hot halves [G][max_nodes] | cold halves [G][max_nodes]
static inline MctsHot *hot_of(void *ws, int64_t G) {
    return reinterpret_cast<MctsHot *>(reinterpret_cast<unsigned char *>(ws) + mcts_nodes_offset(G));
}
static inline MctsCold *cold_of(void *ws, int64_t G, int64_t max_nodes) {
    return reinterpret_cast<MctsCold *>(reinterpret_cast<unsigned char *>(ws) + mcts_nodes_offset(G) + (size_t)G * (size_t)max_nodes * sizeof(MctsHot));
}

extern "C" int aq_mcts_reset(void *ws, const AqState *roots, int64_t G, int64_t max_nodes, void *stream) {
    if (G <= 0 || max_nodes < 1 || !ws || !roots) return aq_set_error(AQ_ERR_ARG, "aq_mcts_reset");
    mcts_reset_kernel<<<(unsigned)((G + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        games_of(ws), hot_of(ws, G), cold_of(ws, G, max_nodes), roots, G, max_nodes);
    return aq_check_launch("aq_mcts_reset");
}

extern "C" int aq_mcts_select(void *ws, int64_t G, int64_t max_nodes, float c_puct, AqState *leaf_states,
                              int32_t *leaf_kind, void *stream) {
    if (G <= 0 || !ws || !leaf_states || !leaf_kind) return aq_set_error(AQ_ERR_ARG, "aq_mcts_select");
    mcts_select_kernel<<<(unsigned)((G + 3) / 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        games_of(ws), hot_of(ws, G), cold_of(ws, G, max_nodes), G, max_nodes, c_puct, leaf_states, leaf_kind);
    return aq_check_launch("aq_mcts_select");
}

extern "C" int aq_mcts_expand_backup(void *ws, int64_t G, int64_t max_nodes, const float *priors, const float *values,
                                     const uint32_t *mask, const uint8_t *pawn, void *stream) {
    if (G <= 0 || !ws || !priors || !values || !mask || !pawn) return aq_set_error(AQ_ERR_ARG, "aq_mcts_expand_backup");
    mcts_expand_backup_kernel<<<(unsigned)((G + 3) / 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        games_of(ws), hot_of(ws, G), cold_of(ws, G, max_nodes), G, max_nodes, priors, values, mask, pawn);
    return aq_check_launch("aq_mcts_expand_backup");
}

extern "C" int aq_mcts_expand_select(void *ws, int64_t G, int64_t max_nodes, const float *priors, const float *values, const uint32_t *mask,
                                     const uint8_t *pawn, float c_puct, AqState *leaf_states, int32_t *leaf_kind, void *stream) {
    if (G <= 0 || !ws || !priors || !values || !mask || !pawn || !leaf_states || !leaf_kind) return aq_set_error(AQ_ERR_ARG, "aq_mcts_expand_select");
    mcts_expand_select_kernel<<<(unsigned)((G + 3) / 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        games_of(ws), hot_of(ws, G), cold_of(ws, G, max_nodes), G, max_nodes, priors, values, mask, pawn, c_puct, leaf_states, leaf_kind);
    return aq_check_launch("aq_mcts_expand_select");
}

extern "C" int aq_mcts_root_counts(void *ws, int64_t G, int64_t max_nodes, int32_t *counts, int16_t *actions,
                                   int16_t *n_children, int32_t *overflow, void *stream) {
    if (G <= 0 || !ws || !counts || !actions || !n_children) return aq_set_error(AQ_ERR_ARG, "aq_mcts_root_counts");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (overflow) {
        cudaError_t e = cudaMemsetAsync(overflow, 0, sizeof(int32_t), st);
        if (e != cudaSuccess) return aq_set_error((int)e, "aq_mcts_root_counts");
    }
    mcts_root_counts_kernel<<<(unsigned)((G + 3) / 4), 128, 0, st>>>(games_of(ws), hot_of(ws, G), cold_of(ws, G, max_nodes), G, max_nodes, counts,
                                                                     actions, n_children, overflow);
    return aq_check_launch("aq_mcts_root_counts");
}

extern "C" int64_t aq_selfplay_ws_bytes(int64_t G) {
    return (int64_t)(((size_t)(G > 0 ? G : 0) * sizeof(AqState) + 255) & ~(size_t)255) + (G > 0 ? G : 0);
}

extern "C" int aq_selfplay_advance(const AqState *states, const int32_t *counts, const int16_t *actions, const int16_t *n_children,
                                   const int64_t *game_id, int64_t G, double temperature, uint64_t seed, int32_t ply, void *policy,
                                   int policy_f64, int16_t *chosen, AqState *next_states, int64_t *next_game_id, uint8_t *final_flags,
                                   int64_t *final_plies, int32_t *alive_count, void *workspace, void *stream) {
    if (G <= 0 || !states || !counts || !actions || !n_children || !game_id || !policy || !next_states || !next_game_id || !final_flags ||
        !final_plies || !alive_count || !workspace || !(temperature >= 0.0))
        return aq_set_error(AQ_ERR_ARG, "aq_selfplay_advance");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    AqState *moved = reinterpret_cast<AqState *>(workspace);
    uint8_t *term = reinterpret_cast<uint8_t *>(workspace) + (((size_t)G * sizeof(AqState) + 255) & ~(size_t)255);
    const int greedy = temperature == 0.0;
    const double inv_t = greedy ? 1.0 : 1.0 / temperature;
    const unsigned grid = (unsigned)((G + 3) / 4);
    if (policy_f64)
        selfplay_move_kernel<double><<<grid, 128, 0, st>>>(states, counts, actions, n_children, game_id, G, inv_t, greedy, seed, ply,
                                                          reinterpret_cast<double *>(policy), chosen, moved, term);
    else
        selfplay_move_kernel<float><<<grid, 128, 0, st>>>(states, counts, actions, n_children, game_id, G, inv_t, greedy, seed, ply,
                                                         reinterpret_cast<float *>(policy), chosen, moved, term);
    int rc = aq_check_launch("aq_selfplay_advance(move)");
    if (rc) return rc;
    selfplay_compact_kernel<<<1, 1024, 0, st>>>(moved, term, game_id, G, ply, next_states, next_game_id, final_flags, final_plies, alive_count);
    return aq_check_launch("aq_selfplay_advance(compact)");
}
