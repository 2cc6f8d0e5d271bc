/*
 * quoridor_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the integer half of the AlphaQuoridorGNN hot path
 * (reference: /root/reference/game_logic.py).  It is the checker the CUDA kernels are
 * compared against; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product never calls it.
 *
 * Pinned: tests/test_oracle_golden.py checks every function below against fixtures
 * produced by the UNMODIFIED reference (tests/golden/make_golden.py, run in the build
 * container where /root/reference exists): G1, KA1..KA13 and 30k+ trajectory positions.
 *
 * Each function cites the reference lines it follows.  The algorithm is kept literal
 * (queue BFS, board rotation for the opponent's search, touch-count gate) on purpose:
 * the CUDA path uses bitboards and an un-rotated opponent search, so agreement between
 * the two is a real check and not the same code twice.
 *
 * State row layout shared with the tests ("row68"): uint8[68] =
 *   [player_pos, player_walls, enemy_pos (enemy's own frame), enemy_walls, walls[64]]
 * walls[s] in {0, 1 = horizontal, 2 = vertical}; for boards smaller than 9x9 only the
 * first (N-1)^2 wall entries are used.  plies_played travels separately (int16).
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OQ_MAXN 9
#define OQ_MAXSLOTS 64
#define OQ_MAXACT 136 /* 5 pawn moves + 128 wall actions, rounded up */

typedef struct {
    int N;
    int ppos, pwalls; /* State.player  (game_logic.py:20) */
    int epos, ewalls; /* State.enemy   (game_logic.py:21), position in the enemy's frame */
    int plies;        /* State.plies_played */
    uint8_t walls[OQ_MAXSLOTS];
} oq_state;

/* ---- is_wall_blocking, game_logic.py:145-167 ------------------------------------- */
static int oq_blocked(const uint8_t *w, int N, int x, int y, int nx, int ny) {
    int M = N - 1;
    if (nx > x) { /* down: horizontal wall in the slot to the bottom-right or bottom-left */
        int br = (y < N - 1) ? (w[x * M + y] == 1) : 0;
        int bl = (y > 0) ? (w[x * M + y - 1] == 1) : 0;
        return br || bl;
    }
    if (nx < x) { /* up */
        int tr = (y < N - 1) ? (w[(x - 1) * M + y] == 1) : 0;
        int tl = (y > 0) ? (w[(x - 1) * M + y - 1] == 1) : 0;
        return tr || tl;
    }
    if (ny > y) { /* right: vertical wall below-right or above-right */
        int br = (x < N - 1) ? (w[x * M + y] == 2) : 0;
        int tr = (x > 0) ? (w[(x - 1) * M + y] == 2) : 0;
        return br || tr;
    }
    if (ny < y) { /* left */
        int bl = (x < N - 1) ? (w[x * M + (y - 1)] == 2) : 0;
        int tl = (x > 0) ? (w[(x - 1) * M + (y - 1)] == 2) : 0;
        return bl || tl;
    }
    return 0;
}

static int oq_inside(int N, int x, int y) { return x >= 0 && x < N && y >= 0 && y < N; }

/* ---- legal_actions_pos, game_logic.py:120-192 ------------------------------------ */
/* MOVEMENT_DIRECTIONS order U, D, L, R (game_logic.py:11). Returns the count. */
int oq_legal_actions_pos(const oq_state *s, int pos, int *out) {
    static const int DX[4] = {-1, 1, 0, 0};
    static const int DY[4] = {0, 0, -1, 1};
    int N = s->N, n = 0;
    const uint8_t *w = s->walls;
    int x = pos / N, y = pos % N;
    int e = (N * N - 1) - s->epos; /* enemy square in the mover's frame, game_logic.py:136 */
    int ex = e / N, ey = e % N;
    for (int d = 0; d < 4; ++d) {
        int dx = DX[d], dy = DY[d];
        int nx = x + dx, ny = y + dy;
        if (!oq_inside(N, nx, ny)) continue;
        if (oq_blocked(w, N, x, y, nx, ny)) continue;
        if (nx == ex && ny == ey) {
            int jx = nx + dx, jy = ny + dy; /* straight jump first, game_logic.py:175-177 */
            if (oq_inside(N, jx, jy) && !oq_blocked(w, N, nx, ny, jx, jy)) {
                out[n++] = jx * N + jy;
            } else if (dx != 0) { /* vertical approach: left then right, :179-183 */
                if (oq_inside(N, nx, ny - 1) && !oq_blocked(w, N, nx, ny, nx, ny - 1)) out[n++] = nx * N + ny - 1;
                if (oq_inside(N, nx, ny + 1) && !oq_blocked(w, N, nx, ny, nx, ny + 1)) out[n++] = nx * N + ny + 1;
            } else { /* horizontal approach: up then down, :184-188 */
                if (oq_inside(N, nx - 1, ny) && !oq_blocked(w, N, nx, ny, nx - 1, ny)) out[n++] = (nx - 1) * N + ny;
                if (oq_inside(N, nx + 1, ny) && !oq_blocked(w, N, nx, ny, nx + 1, ny)) out[n++] = (nx + 1) * N + ny;
            }
        } else {
            out[n++] = nx * N + ny;
        }
    }
    return n;
}

/* ---- can_place_wall, game_logic.py:199-223 --------------------------------------- */
static int oq_can_place(const uint8_t *w, int N, int o, int pos) {
    int M = N - 1;
    if (w[pos] != 0) return 0;
    int x = pos / M, y = pos % M;
    if (o == 1) {
        if (y > 0 && w[pos - 1] == 1) return 0;
        if (y < N - 2 && w[pos + 1] == 1) return 0;
    } else {
        if (x > 0 && w[pos - M] == 2) return 0;
        if (x < N - 2 && w[pos + M] == 2) return 0;
    }
    return 1;
}

/* ---- is_goal_possibly_blocked, game_logic.py:227-307 ------------------------------
 * The if/elif chains in the reference all set the same flag, so each end reduces to a
 * guarded OR; the guards reproduce the index-range conditions of the reference. */
static int oq_gate(const uint8_t *w, int N, int o, int pos) {
    int M = N - 1;
    int x = pos / M, y = pos % M;
    int a, mid, b;
    if (o == 1) {
        a = (y == 0) ||
            (y > 0 && (w[pos - 1] == 2 || (x > 0 && w[pos - M - 1] == 2) || (x < N - 2 && w[pos + M - 1] == 2))) ||
            (y > 1 && w[pos - 2] == 1);
        mid = (x > 0 && w[pos - M] == 2) || (x < N - 2 && w[pos + M] == 2);
        b = (y == N - 2) ||
            (y < N - 2 && (w[pos + 1] == 2 || (x > 0 && w[pos - M + 1] == 2) || (x < N - 2 && w[pos + M + 1] == 2))) ||
            (y < N - 3 && w[pos + 2] == 1);
    } else {
        a = (x == 0) ||
            (x > 0 && (w[pos - M] == 1 || (y > 0 && w[pos - M - 1] == 1) || (y < N - 2 && w[pos - M + 1] == 1))) ||
            (x > 1 && w[pos - 2 * M] == 2);
        mid = (y > 0 && w[pos - 1] == 1) || (y < N - 2 && w[pos + 1] == 1);
        b = (x == N - 2) ||
            (x < N - 2 && (w[pos + M] == 1 || (y > 0 && w[pos + M - 1] == 1) || (y < N - 2 && w[pos + M + 1] == 1))) ||
            (x < N - 3 && w[pos + 2 * M] == 2);
    }
    return (a + mid + b) >= 2;
}

/* ---- State.next + rotate_walls, game_logic.py:359-391 ----------------------------- */
void oq_next(const oq_state *s, int action, oq_state *out) {
    int N = s->N, M = N - 1, S = M * M;
    oq_state t = *s;
    t.plies = s->plies + 1;
    if (action < N * N) {
        t.ppos = action;
    } else if (action < N * N + S) {
        t.walls[action - N * N] = 1;
        t.pwalls -= 1;
    } else {
        t.walls[action - N * N - S] = 2;
        t.pwalls -= 1;
    }
    uint8_t r[OQ_MAXSLOTS];
    memset(r, 0, sizeof r);
    for (int i = 0; i < S; ++i) r[i] = t.walls[S - 1 - i];
    memcpy(t.walls, r, sizeof r);
    int p = t.ppos, pw = t.pwalls;
    t.ppos = t.epos; t.pwalls = t.ewalls;
    t.epos = p; t.ewalls = pw;
    *out = t;
}

/* ---- bfs, game_logic.py:309-324 --------------------------------------------------- */
static int oq_bfs(const oq_state *s) {
    int N = s->N;
    uint8_t seen[OQ_MAXN * OQ_MAXN];
    int queue[OQ_MAXN * OQ_MAXN];
    int head = 0, tail = 0, nb[8];
    memset(seen, 0, sizeof seen);
    seen[s->ppos] = 1;
    queue[tail++] = s->ppos;
    while (head < tail) {
        int p = queue[head++];
        if (p / N == 0) return 1;
        int k = oq_legal_actions_pos(s, p, nb);
        for (int i = 0; i < k; ++i)
            if (!seen[nb[i]]) { seen[nb[i]] = 1; queue[tail++] = nb[i]; }
    }
    return 0;
}

/* ---- can_reach_goal, game_logic.py:225-348 ---------------------------------------- */
static int oq_can_reach_goal(const oq_state *s, int o, int pos) {
    int N = s->N, M = N - 1;
    if (!oq_gate(s->walls, N, o, pos)) return 1; /* BFS skipped, game_logic.py:327-328 */
    oq_state ps = *s;
    ps.walls[pos] = (uint8_t)o;
    int mover_ok = oq_bfs(&ps);
    int action = pos + (o == 1 ? N * N : N * N + M * M);
    oq_state es;
    oq_next(&ps, action, &es); /* rotated board, roles swapped, game_logic.py:344 */
    int enemy_ok = oq_bfs(&es);
    return mover_ok && enemy_ok;
}

/* ---- legal_actions (+ legal_actions_wall), game_logic.py:103-117, 350-357 --------- */
int oq_legal_actions(const oq_state *s, int *out) {
    int N = s->N, M = N - 1, S = M * M;
    int n = oq_legal_actions_pos(s, s->ppos, out);
    if (s->pwalls > 0) {
        for (int pos = 0; pos < S; ++pos) {
            if (oq_can_place(s->walls, N, 1, pos) && oq_can_reach_goal(s, 1, pos)) out[n++] = N * N + pos;
            if (oq_can_place(s->walls, N, 2, pos) && oq_can_reach_goal(s, 2, pos)) out[n++] = N * N + S + pos;
        }
    }
    return n;
}

int oq_legal_actions_wall(const oq_state *s, int pos, int *out) {
    int N = s->N, M = N - 1, S = M * M, n = 0;
    if (oq_can_place(s->walls, N, 1, pos) && oq_can_reach_goal(s, 1, pos)) out[n++] = N * N + pos;
    if (oq_can_place(s->walls, N, 2, pos) && oq_can_reach_goal(s, 2, pos)) out[n++] = N * N + S + pos;
    return n;
}

/* is_lose / is_draw, game_logic.py:43-50 */
int oq_is_lose(const oq_state *s) { return s->epos / s->N == 0; }
int oq_is_draw(const oq_state *s, int plies_for_draw) { return s->plies >= plies_for_draw; }

/* ---- row68 <-> struct ------------------------------------------------------------- */
static void oq_load(const uint8_t *row, int plies, int N, oq_state *s) {
    memset(s, 0, sizeof *s);
    s->N = N;
    s->ppos = row[0]; s->pwalls = row[1]; s->epos = row[2]; s->ewalls = row[3];
    s->plies = plies;
    memcpy(s->walls, row + 4, (size_t)((N - 1) * (N - 1)));
}
static void oq_store(const oq_state *s, uint8_t *row, int16_t *plies) {
    row[0] = (uint8_t)s->ppos; row[1] = (uint8_t)s->pwalls;
    row[2] = (uint8_t)s->epos; row[3] = (uint8_t)s->ewalls;
    memcpy(row + 4, s->walls, OQ_MAXSLOTS);
    if (plies) *plies = (int16_t)s->plies;
}

/* ---- batch entry points (ctypes) --------------------------------------------------
 * actions:  int16[M, OQ_MAXACT] ordered exactly like State.legal_actions(), -1 padded
 * nactions: int16[M]
 * mask:     uint32[M, 8]  bit a of the 256-bit little-endian word string set iff action a legal
 * pawn:     uint8[M, 8]   = [n_pawn, p0..p4 (0xFF padded), 0, 0]  ordered pawn moves
 */
void oq_legal_actions_batch(const uint8_t *rows, const int16_t *plies, long long M, int N,
                            int16_t *actions, int16_t *nactions, uint32_t *mask, uint8_t *pawn,
                            int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < M; ++i) {
        oq_state s;
        int buf[OQ_MAXACT];
        oq_load(rows + 68 * i, plies ? plies[i] : 0, N, &s);
        int n = oq_legal_actions(&s, buf);
        if (nactions) nactions[i] = (int16_t)n;
        if (actions) {
            for (int k = 0; k < OQ_MAXACT; ++k) actions[i * OQ_MAXACT + k] = (int16_t)(k < n ? buf[k] : -1);
        }
        if (mask) {
            uint32_t *m = mask + 8 * i;
            for (int k = 0; k < 8; ++k) m[k] = 0;
            for (int k = 0; k < n; ++k) m[buf[k] >> 5] |= 1u << (buf[k] & 31);
        }
        if (pawn) {
            uint8_t *p = pawn + 8 * i;
            int np = 0;
            for (int k = 0; k < 8; ++k) p[k] = (k >= 1 && k <= 5) ? 0xFF : 0;
            while (np < n && buf[np] < N * N) { p[1 + np] = (uint8_t)buf[np]; ++np; }
            p[0] = (uint8_t)np;
        }
    }
}

int oq_legal_actions_pos_row(const uint8_t *row, int N, int pos, int *out) {
    oq_state s;
    oq_load(row, 0, N, &s);
    return oq_legal_actions_pos(&s, pos, out);
}

int oq_legal_actions_wall_row(const uint8_t *row, int N, int pos, int *out) {
    oq_state s;
    oq_load(row, 0, N, &s);
    return oq_legal_actions_wall(&s, pos, out);
}

/* next(): rows_out uint8[M,68], plies_out int16[M], flags_out uint8[M] bit0 = is_lose of
 * the successor, bit1 = is_draw of the successor. */
void oq_next_batch(const uint8_t *rows, const int16_t *plies, const int16_t *actions, long long M,
                   int N, int plies_for_draw, uint8_t *rows_out, int16_t *plies_out,
                   uint8_t *flags_out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < M; ++i) {
        oq_state s, t;
        oq_load(rows + 68 * i, plies[i], N, &s);
        oq_next(&s, actions[i], &t);
        oq_store(&t, rows_out + 68 * i, plies_out + i);
        if (flags_out) flags_out[i] = (uint8_t)(oq_is_lose(&t) | (oq_is_draw(&t, plies_for_draw) << 1));
    }
}

/* ---- graph + features (derived) ----------------------
 * open[v] bit k set iff direction k of MOVEMENT_DIRECTIONS (U,D,L,R) leads to an
 * in-board square and is not wall-blocked (game_logic.py:145-167); pawns ignored.
 * planes: float32[6,N,N] following State.pieces_array (game_logic.py:56-93) /
 * CNNNetwork.preprocess_input (pv_network_cnn.py:97-112). */
void oq_open_mask_batch(const uint8_t *rows, long long M, int N, uint8_t *open) {
    static const int DX[4] = {-1, 1, 0, 0};
    static const int DY[4] = {0, 0, -1, 1};
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < M; ++i) {
        const uint8_t *w = rows + 68 * i + 4;
        for (int v = 0; v < N * N; ++v) {
            int x = v / N, y = v % N, m = 0;
            for (int d = 0; d < 4; ++d) {
                int nx = x + DX[d], ny = y + DY[d];
                if (oq_inside(N, nx, ny) && !oq_blocked(w, N, x, y, nx, ny)) m |= 1 << d;
            }
            open[i * N * N + v] = (uint8_t)m;
        }
    }
}

void oq_planes_batch(const uint8_t *rows, long long M, int N, float *planes) {
    int V = N * N, Mw = N - 1;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < M; ++i) {
        const uint8_t *r = rows + 68 * i;
        float *p = planes + i * 6 * V;
        for (int k = 0; k < 6 * V; ++k) p[k] = 0.f;
        p[0 * V + r[0]] = 1.f;
        for (int v = 0; v < V; ++v) p[1 * V + v] = (float)r[1];
        p[2 * V + r[2]] = 1.f; /* enemy square in the enemy's own frame (quirk KA11) */
        for (int v = 0; v < V; ++v) p[3 * V + v] = (float)r[3];
        for (int sidx = 0; sidx < Mw * Mw; ++sidx) {
            int tile = N * (sidx / Mw) + (sidx % Mw);
            if (r[4 + sidx] == 1) p[4 * V + tile] = 1.f;
            else if (r[4 + sidx] == 2) p[5 * V + tile] = 1.f;
        }
    }
}

/* ---- agents.py: heuristic evaluation and alpha-beta ---- */
/* shortest_path_bfs, agents.py:27-41: queue of (position, depth) pairs; -1 if no path. */
static int oq_shortest_path(const oq_state *s) {
    int N = s->N;
    uint8_t seen[OQ_MAXN * OQ_MAXN];
    int qpos[OQ_MAXN * OQ_MAXN], qdepth[OQ_MAXN * OQ_MAXN];
    int head = 0, tail = 0, nb[8];
    memset(seen, 0, sizeof seen);
    seen[s->ppos] = 1;
    qpos[tail] = s->ppos; qdepth[tail++] = 0;
    while (head < tail) {
        int p = qpos[head], d = qdepth[head++];
        if (p / N == 0) return d;
        int k = oq_legal_actions_pos(s, p, nb);
        for (int i = 0; i < k; ++i)
            if (!seen[nb[i]]) { seen[nb[i]] = 1; qpos[tail] = nb[i]; qdepth[tail++] = d + 1; }
    }
    return -1;
}

/* shortest_path_diff, agents.py:43-52: the enemy's search runs on the rotated board with the roles swapped. */
static void oq_shortest_paths(const oq_state *s, int *dp, int *de) {
    int S = (s->N - 1) * (s->N - 1);
    *dp = oq_shortest_path(s);
    oq_state t = *s;
    for (int i = 0; i < S; ++i) t.walls[i] = s->walls[S - 1 - i]; /* rotate_walls, game_logic.py:359-364 */
    t.ppos = s->epos; t.pwalls = s->ewalls;
    t.epos = s->ppos; t.ewalls = s->pwalls;
    *de = oq_shortest_path(&t);
}

static double oq_heuristic(const oq_state *s, int plies_for_draw, int num_walls) {
    int dp, de;
    oq_shortest_paths(s, &dp, &de);
    return (double)(de - dp) / (double)(plies_for_draw / 2 - num_walls); /* agents.py:11,52 */
}

void oq_heuristic_batch(const uint8_t *rows, long long M, int N, int plies_for_draw, int num_walls,
                        int16_t *dist, double *heur) {
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < M; ++i) {
        oq_state s;
        int dp, de;
        oq_load(rows + 68 * i, 0, N, &s);
        oq_shortest_paths(&s, &dp, &de);
        if (dist) { dist[2 * i] = (int16_t)dp; dist[2 * i + 1] = (int16_t)de; }
        if (heur) heur[i] = (double)(de - dp) / (double)(plies_for_draw / 2 - num_walls);
    }
}

/* alpha_beta, agents.py:58-86 (fail-hard, returns alpha), kept literal including the pruning. */
static double oq_alpha_beta(const oq_state *s, double alpha, double beta, int depth, int plies_for_draw, int num_walls) {
    if (depth == 0 || oq_is_lose(s) || oq_is_draw(s, plies_for_draw)) {
        if (oq_is_lose(s)) return -1.0;
        if (oq_is_draw(s, plies_for_draw)) return 0.0;
        return oq_heuristic(s, plies_for_draw, num_walls);
    }
    int acts[OQ_MAXACT];
    int n = oq_legal_actions(s, acts);
    for (int i = 0; i < n; ++i) {
        oq_state t;
        oq_next(s, acts[i], &t);
        double score = -oq_alpha_beta(&t, -beta, -alpha, depth - 1, plies_for_draw, num_walls);
        if (score > alpha) alpha = score;
        if (alpha >= beta) return alpha;
    }
    return alpha;
}

/* alpha_beta_action, agents.py:90-107. scores (may be NULL): the score the loop saw for every root action. */
int oq_alpha_beta_action(const uint8_t *row, int plies, int N, int max_depth, int plies_for_draw, int num_walls,
                         double *scores) {
    oq_state s;
    oq_load(row, plies, N, &s);
    int acts[OQ_MAXACT];
    int n = oq_legal_actions(&s, acts), best = -1;
    double alpha = -1.0 / 0.0;
    for (int i = 0; i < n; ++i) {
        oq_state t;
        oq_next(&s, acts[i], &t);
        double score = -oq_alpha_beta(&t, -1.0 / 0.0, -alpha, max_depth, plies_for_draw, num_walls);
        if (scores) scores[i] = score;
        if (score > alpha) { best = acts[i]; alpha = score; }
    }
    return best;
}

int oq_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
