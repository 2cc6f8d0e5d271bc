// Shared device helpers: packed state, 81-bit square bitboards, 64-bit wall-slot bitboards.
// Board geometry follows game_logic.py: squares row-major idx = 9*row + col, wall slot
// s = 8*row + col is the 2x2 block whose top-left tile is (row, col)  (game_logic.py:17-19).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/aqgnn.h"

typedef unsigned __int128 u128;
typedef unsigned long long u64;

#define AQ_DEV __host__ __device__ __forceinline__

int aq_set_error(int code, const char *what);
int aq_check_launch(const char *what);

// Programmatic dependent launch (PDL).  A kernel launched with aq_launch_pdl may start while its predecessor in the stream is still
// draining; it must call aq_pdl_wait() before it touches anything the predecessor (or anything earlier in the stream) writes, and may
// run a prologue that reads only long-lived inputs (parameters) before that.  A predecessor that calls aq_pdl_trigger() at its start
// lets the dependent grid be scheduled as soon as SM resources free up; without the trigger the dependent starts when the
// predecessor's blocks have exited, as with a normal launch.  Both device calls are no-ops for normally launched kernels.
// RULE: what the predecessor wrote must be read with coherent loads (__ldcg / plain loads through a non-restrict pointer) after
// aq_pdl_wait().  A load through a `const __restrict__` pointer or __ldg() is an invariant (LDG.CONSTANT) load to the compiler and
// gets hoisted ABOVE the wait -- seen in the SASS of legal_search_kernel, which then read the previous call's task count.
#ifdef __CUDACC__
__device__ __forceinline__ void aq_pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void aq_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t aq_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#endif

namespace aq {

AQ_DEV u128 bit81(int sq) { return (u128)1 << sq; }
AQ_DEV constexpr u128 make128(u64 hi, u64 lo) { return ((u128)hi << 64) | lo; }

// rows r has bits 9r..9r+8
constexpr u128 kFull = (((u128)1) << 81) - 1;
constexpr u128 kRow0 = (u128)0x1FF;
constexpr u128 kRow8 = (u128)0x1FF << 72;
// column 0: bits 0,9,18,...,72
constexpr u128 kCol0 = ((u128)1) | ((u128)1 << 9) | ((u128)1 << 18) | ((u128)1 << 27) | ((u128)1 << 36) |
                       ((u128)1 << 45) | ((u128)1 << 54) | ((u128)1 << 63) | ((u128)1 << 72);
constexpr u128 kCol8 = kCol0 << 8;

constexpr u64 kC0 = 0x0101010101010101ull;  // wall-slot column 0
constexpr u64 kC7 = 0x8080808080808080ull;  // wall-slot column 7
constexpr u64 kR0 = 0x00000000000000FFull;  // wall-slot row 0
constexpr u64 kR7 = 0xFF00000000000000ull;  // wall-slot row 7

// 8x8 slot board -> 9x9 square board, slot (x,y) -> square (x,y) (its top-left tile)
AQ_DEV u128 expand8to9(u64 b) {
    u128 r = 0;
#pragma unroll
    for (int x = 0; x < 8; ++x) r |= (u128)((b >> (8 * x)) & 0xFFull) << (9 * x);
    return r;
}

// Open-direction bitboards: bit sq of up/down/left/right set iff the pawn move from sq in that
// direction stays on the board and is not wall-blocked (is_wall_blocking, game_logic.py:145-167).
struct Open {
    u128 up, down, left, right;
};

AQ_DEV Open open_from_walls(u64 h, u64 v) {
    const u128 eh = expand8to9(h), ev = expand8to9(v);
    const u128 bdown = eh | (eh << 1);   // H wall in slot (x,y) blocks squares (x,y),(x,y+1) downward
    const u128 bright = ev | (ev << 9);  // V wall in slot (x,y) blocks squares (x,y),(x+1,y) rightward
    Open o;
    o.down = ~bdown & ~kRow8 & kFull;
    o.up = ~(bdown << 9) & ~kRow0 & kFull;
    o.right = ~bright & ~kCol8 & kFull;
    o.left = ~(bright << 1) & ~kCol0 & kFull;
    return o;
}

// remove the edges severed by one extra wall (o = 1 horizontal, 2 vertical) in slot (x,y)
AQ_DEV void add_wall(Open &o, int orient, int slot) {
    const int x = slot >> 3, y = slot & 7;
    const int sq = 9 * x + y;
    if (orient == 1) {
        const u128 m = (u128)3 << sq;
        o.down &= ~m;
        o.up &= ~(m << 9);
    } else {
        const u128 m = ((u128)1 | ((u128)1 << 9)) << sq;
        o.right &= ~m;
        o.left &= ~(m << 1);
    }
}

AQ_DEV bool has(u128 b, int sq) { return (unsigned)((b >> sq) & 1) != 0u; }

// Targets of a jump over the pawn on `ob` when it is approached moving in direction d
// (0=U,1=D,2=L,3=R): straight if open, else the two perpendicular squares that are open
// (game_logic.py:174-188).  Returned as a bitboard (order is irrelevant for reachability).
AQ_DEV u128 jump_targets(const Open &o, int ob, int d) {
    u128 t = 0;
    if (d == 0) {
        if (has(o.up, ob)) t = bit81(ob - 9);
        else { if (has(o.left, ob)) t |= bit81(ob - 1); if (has(o.right, ob)) t |= bit81(ob + 1); }
    } else if (d == 1) {
        if (has(o.down, ob)) t = bit81(ob + 9);
        else { if (has(o.left, ob)) t |= bit81(ob - 1); if (has(o.right, ob)) t |= bit81(ob + 1); }
    } else if (d == 2) {
        if (has(o.left, ob)) t = bit81(ob - 1);
        else { if (has(o.up, ob)) t |= bit81(ob - 9); if (has(o.down, ob)) t |= bit81(ob + 9); }
    } else {
        if (has(o.right, ob)) t = bit81(ob + 1);
        else { if (has(o.up, ob)) t |= bit81(ob - 9); if (has(o.down, ob)) t |= bit81(ob + 9); }
    }
    return t;
}

// bfs() of game_logic.py:309-324 as a bitboard flood fill with the pawn rules of
// legal_actions_pos applied at every visited square: the obstacle square `ob` is never entered;
// a square adjacent to it (edge open) reaches the jump targets instead.  Returns true iff a
// square of `goal` is reachable from `start`.
AQ_DEV bool reaches(const Open &o, int start, int ob, u128 goal) {
    // source squares from which a move in direction d lands on the obstacle
    const u128 srcU = has(o.down, ob) ? bit81(ob + 9) : (u128)0;  // below, moving up
    const u128 srcD = has(o.up, ob) ? bit81(ob - 9) : (u128)0;
    const u128 srcL = has(o.right, ob) ? bit81(ob + 1) : (u128)0;  // right of it, moving left
    const u128 srcR = has(o.left, ob) ? bit81(ob - 1) : (u128)0;
    const u128 jU = jump_targets(o, ob, 0), jD = jump_targets(o, ob, 1);
    const u128 jL = jump_targets(o, ob, 2), jR = jump_targets(o, ob, 3);
    const u128 notob = ~bit81(ob);
    u128 reach = bit81(start);
    while (true) {
        if (reach & goal) return true;
        u128 nb = ((reach & o.up) >> 9) | ((reach & o.down) << 9) | ((reach & o.left) >> 1) | ((reach & o.right) << 1);
        nb &= notob;
        if (reach & srcU) nb |= jU;
        if (reach & srcD) nb |= jD;
        if (reach & srcL) nb |= jL;
        if (reach & srcR) nb |= jR;
        const u128 nxt = reach | nb;
        if (nxt == reach) return false;
        reach = nxt;
    }
}

// shortest_path_bfs of agents.py:27-41: number of pawn steps (a jump is one step) from `start` to the nearest
// square of `goal` under the same pawn rules; every iteration of the flood fill is one BFS layer.  -1 if no
// path exists (agents.py:41).
AQ_DEV int path_length(const Open &o, int start, int ob, u128 goal) {
    const u128 srcU = has(o.down, ob) ? bit81(ob + 9) : (u128)0;
    const u128 srcD = has(o.up, ob) ? bit81(ob - 9) : (u128)0;
    const u128 srcL = has(o.right, ob) ? bit81(ob + 1) : (u128)0;
    const u128 srcR = has(o.left, ob) ? bit81(ob - 1) : (u128)0;
    const u128 jU = jump_targets(o, ob, 0), jD = jump_targets(o, ob, 1);
    const u128 jL = jump_targets(o, ob, 2), jR = jump_targets(o, ob, 3);
    const u128 notob = ~bit81(ob);
    u128 reach = bit81(start);
    for (int depth = 0;; ++depth) {
        if (reach & goal) return depth;
        u128 nb = ((reach & o.up) >> 9) | ((reach & o.down) << 9) | ((reach & o.left) >> 1) | ((reach & o.right) << 1);
        nb &= notob;
        if (reach & srcU) nb |= jU;
        if (reach & srcD) nb |= jD;
        if (reach & srcL) nb |= jL;
        if (reach & srcR) nb |= jR;
        const u128 nxt = reach | nb;
        if (nxt == reach) return -1;
        reach = nxt;
    }
}

// ---- path witness --------------------------------------------------------------------------------
// One concrete path start -> goal (with the pawn rules) found by the same flood fill, returned as the
// set of wall slots whose H / V wall would sever one of the unit edges the path uses.  A candidate
// wall outside these sets leaves the path intact, so the goal stays reachable and no search is needed:
// walls only ADD blocks, every "edge open" condition the path relies on stays true unless that edge is
// severed, and every "straight jump blocked" condition stays true.  (The converse is not used: a
// candidate that cuts the witness gets a real search.  If no path exists without the candidate, all
// candidates are searched -- a wall behind the other pawn can create diagonal jumps.)
struct PathCuts {
    u64 cutH, cutV;
    int exists;
};

AQ_DEV int lowest_bit(u128 b) {
    const u64 lo = (u64)b, hi = (u64)(b >> 64);
#ifdef __CUDA_ARCH__
    return lo ? __ffsll((long long)lo) - 1 : 64 + __ffsll((long long)hi) - 1;
#else
    return lo ? __builtin_ctzll(lo) : 64 + __builtin_ctzll(hi);
#endif
}

// unit edge between adjacent squares a and b -> slots whose wall severs it
AQ_DEV void add_edge_cut(PathCuts &pc, int a, int b) {
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    const int r = lo / 9, c = lo % 9;
    if (hi - lo == 9) {  // vertical edge (r,c)-(r+1,c): H walls in slots (r,c) and (r,c-1)
        if (c < 8) pc.cutH |= 1ull << (8 * r + c);
        if (c > 0) pc.cutH |= 1ull << (8 * r + c - 1);
    } else {             // horizontal edge (r,c)-(r,c+1): V walls in slots (r,c) and (r-1,c)
        if (r < 8) pc.cutV |= 1ull << (8 * r + c);
        if (r > 0) pc.cutV |= 1ull << (8 * (r - 1) + c);
    }
}

AQ_DEV PathCuts find_path_cuts(const Open &o, int start, int ob, u128 goal) {
    const u128 srcU = has(o.down, ob) ? bit81(ob + 9) : (u128)0;
    const u128 srcD = has(o.up, ob) ? bit81(ob - 9) : (u128)0;
    const u128 srcL = has(o.right, ob) ? bit81(ob + 1) : (u128)0;
    const u128 srcR = has(o.left, ob) ? bit81(ob - 1) : (u128)0;
    const u128 jU = jump_targets(o, ob, 0), jD = jump_targets(o, ob, 1);
    const u128 jL = jump_targets(o, ob, 2), jR = jump_targets(o, ob, 3);
    const u128 notob = ~bit81(ob);
    u128 reach = bit81(start);
    // squares first reached by a plain move in direction U/D/L/R, or by a jump approached in that direction
    u128 byU = 0, byD = 0, byL = 0, byR = 0, jbU = 0, jbD = 0, jbL = 0, jbR = 0;
    PathCuts pc;
    pc.cutH = 0; pc.cutV = 0; pc.exists = 0;
    while (!(reach & goal)) {
        u128 acc = reach, n;
        n = ((reach & o.up) >> 9) & notob & ~acc;    byU |= n; acc |= n;
        n = ((reach & o.down) << 9) & notob & ~acc;  byD |= n; acc |= n;
        n = ((reach & o.left) >> 1) & notob & ~acc;  byL |= n; acc |= n;
        n = ((reach & o.right) << 1) & notob & ~acc; byR |= n; acc |= n;
        if (reach & srcU) { n = jU & ~acc; jbU |= n; acc |= n; }
        if (reach & srcD) { n = jD & ~acc; jbD |= n; acc |= n; }
        if (reach & srcL) { n = jL & ~acc; jbL |= n; acc |= n; }
        if (reach & srcR) { n = jR & ~acc; jbR |= n; acc |= n; }
        if (acc == reach) return pc;  // goal unreachable even without a candidate
        reach = acc;
    }
    pc.exists = 1;
    int cur = lowest_bit(reach & goal);
    while (cur != start) {
        int prev;
        if (has(byU, cur)) prev = cur + 9;
        else if (has(byD, cur)) prev = cur - 9;
        else if (has(byL, cur)) prev = cur + 1;
        else if (has(byR, cur)) prev = cur - 1;
        else {  // jump over the pawn on ob: two unit edges, src-ob and ob-cur
            prev = has(jbU, cur) ? ob + 9 : has(jbD, cur) ? ob - 9 : has(jbL, cur) ? ob + 1 : ob - 1;
            add_edge_cut(pc, prev, ob);
            add_edge_cut(pc, ob, cur);
            cur = prev;
            continue;
        }
        add_edge_cut(pc, prev, cur);
        cur = prev;
    }
    return pc;
}

// legal_actions_pos (game_logic.py:120-192) from square p with the enemy pawn on e (mover's
// frame); writes up to 5 squares in the reference order, returns the count.
AQ_DEV int pawn_moves(const Open &o, int p, int e, uint8_t *out) {
    int n = 0;
    const int delta[4] = {-9, 9, -1, 1};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const u128 od = d == 0 ? o.up : d == 1 ? o.down : d == 2 ? o.left : o.right;
        if (!has(od, p)) continue;
        const int q = p + delta[d];
        if (q != e) { out[n++] = (uint8_t)q; continue; }
        if (has(od, q)) { out[n++] = (uint8_t)(q + delta[d]); continue; }
        if (d < 2) {
            if (has(o.left, q)) out[n++] = (uint8_t)(q - 1);
            if (has(o.right, q)) out[n++] = (uint8_t)(q + 1);
        } else {
            if (has(o.up, q)) out[n++] = (uint8_t)(q - 9);
            if (has(o.down, q)) out[n++] = (uint8_t)(q + 9);
        }
    }
    return n;
}

// ---- wall-slot (8x8) neighbourhood shifts: value at slot s taken from its W/E/N/S neighbour --
AQ_DEV u64 fromW(u64 b) { return (b & ~kC7) << 1; }  // s-1, needs col > 0
AQ_DEV u64 fromE(u64 b) { return (b & ~kC0) >> 1; }  // s+1, needs col < 7
AQ_DEV u64 fromN(u64 b) { return b << 8; }           // s-8, needs row > 0
AQ_DEV u64 fromS(u64 b) { return b >> 8; }           // s+8, needs row < 7

struct WallSets {
    u64 freeH, freeV;  // can_place_wall true and the gate does NOT fire -> legal without search
    u64 needH, needV;  // can_place_wall true and the gate fires -> needs both searches
};

// can_place_wall (game_logic.py:199-223) and is_goal_possibly_blocked (227-307) for all 64
// slots and both orientations at once.
AQ_DEV WallSets wall_sets(u64 h, u64 v) {
    const u64 occ = h | v;
    const u64 canH = ~occ & ~fromW(h) & ~fromE(h);
    const u64 canV = ~occ & ~fromN(v) & ~fromS(v);
    // horizontal candidate: left end, middle, right end
    const u64 lH = kC0 | fromW(v) | fromW(fromN(v)) | fromW(fromS(v)) | fromW(fromW(h));
    const u64 mH = fromN(v) | fromS(v);
    const u64 rH = kC7 | fromE(v) | fromE(fromN(v)) | fromE(fromS(v)) | fromE(fromE(h));
    const u64 gateH = (lH & mH) | (lH & rH) | (mH & rH);
    // vertical candidate: top end, middle, bottom end
    const u64 tV = kR0 | fromN(h) | fromN(fromW(h)) | fromN(fromE(h)) | fromN(fromN(v));
    const u64 mV = fromW(h) | fromE(h);
    const u64 bV = kR7 | fromS(h) | fromS(fromW(h)) | fromS(fromE(h)) | fromS(fromS(v));
    const u64 gateV = (tV & mV) | (tV & bV) | (mV & bV);
    WallSets w;
    w.freeH = canH & ~gateH;
    w.needH = canH & gateH;
    w.freeV = canV & ~gateV;
    w.needV = canV & gateV;
    return w;
}

__device__ __forceinline__ AqState load_state(const AqState *p) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint2 b = __ldg(reinterpret_cast<const uint2 *>(p) + 2);
    AqState s;
    s.hwalls = ((u64)a.y << 32) | a.x;
    s.vwalls = ((u64)a.w << 32) | a.z;
    s.ppos = b.x & 0xFF;
    s.pwalls = (b.x >> 8) & 0xFF;
    s.epos = (b.x >> 16) & 0xFF;
    s.ewalls = (b.x >> 24) & 0xFF;
    s.plies = b.y & 0xFFFF;
    s.flags = (b.y >> 16) & 0xFFFF;
    s.reserved = 0;
    return s;
}

// State.next() (game_logic.py:366-391): apply the action, rotate the board 180 degrees
// (rotate_walls = 64-bit bit reversal of each wall bitboard), swap the players, plies + 1.
__device__ __forceinline__ AqState state_after(AqState s, int a) {
    int ppos = s.ppos, pwalls = s.pwalls;
    if (a < AQ_SQUARES) {
        ppos = a;
    } else if (a < AQ_SQUARES + AQ_SLOTS) {
        s.hwalls |= 1ull << (a - AQ_SQUARES);
        pwalls -= 1;
    } else {
        s.vwalls |= 1ull << (a - AQ_SQUARES - AQ_SLOTS);
        pwalls -= 1;
    }
    AqState t;
    t.hwalls = __brevll(s.hwalls);
    t.vwalls = __brevll(s.vwalls);
    t.ppos = s.epos;
    t.pwalls = s.ewalls;
    t.epos = (uint8_t)ppos;
    t.ewalls = (uint8_t)pwalls;
    t.plies = (uint16_t)(s.plies + 1);
    t.flags = 0;
    t.reserved = 0;
    return t;
}
// is_lose (game_logic.py:43-46) | is_draw << 1 (game_logic.py:49-50)
__device__ __forceinline__ int terminal_flags(const AqState &s) {
    return (int)((s.epos / AQ_N) == 0) | ((int)(s.plies >= AQ_PLIES_FOR_DRAW) << 1);
}
__device__ __forceinline__ void store_state(AqState *p, const AqState &s) {
    uint4 x, y;
    x.x = (unsigned)s.hwalls; x.y = (unsigned)(s.hwalls >> 32); x.z = (unsigned)s.vwalls; x.w = (unsigned)(s.vwalls >> 32);
    y.x = (unsigned)s.ppos | ((unsigned)s.pwalls << 8) | ((unsigned)s.epos << 16) | ((unsigned)s.ewalls << 24);
    y.y = (unsigned)s.plies;
    y.z = 0; y.w = 0;
    reinterpret_cast<uint4 *>(p)[0] = x;
    reinterpret_cast<uint4 *>(p)[1] = y;
}

}  // namespace aq
