"""Policy-Value Monte Carlo Tree Search -- drop-in for reference pv_mcts.py, batched over games.

Same public names (PV_EVALUATE_COUNT, pv_mcts_policy, pv_mcts_action, boltzman).  The tree lives on
the GPU (csrc/mcts.cu): G games are searched in lock-step, so the leaf of every game at simulation k
forms one batch for the network (batched leaf evaluation = model.predict for G states at once,
pv_mcts.py:47).  The reference has no virtual loss, no Dirichlet noise and no tree reuse
(pv_mcts.py:81 builds a fresh root per move), so each game's search is exactly the reference's.
"""
from copy import deepcopy

import numpy as np
import torch

from . import _lib
from . import game_logic as gl

# Prepare parameters
PV_EVALUATE_COUNT = 50  # Number of simulations per inference (pv_mcts.py:18)
C_PUCT = 1.25           # pv_mcts.py:71
MAX_CHILDREN = 133      # 5 pawn moves + 128 wall placements
GRAPH_CHUNK = 25        # simulations per launch of the chunk graph (BatchedMCTS.search)


def network_evaluator(model):
    """Evaluator protocol: packed states uint8[G,32] -> dict(priors f32[G,209] (legal-masked,
    renormalised), value f32[G], mask int32[G,8], pawn uint8[G,8])."""
    return lambda packed: model.predict_batch(packed)


class BatchedMCTS:
    """Lock-step PV-MCTS over G independent root states.

    `evaluator` is either a GNNNetwork (leaf evaluation = aq_leaf_eval into preallocated buffers, and the
    whole simulation step -- select, legal mask, GNN trunk, heads, expand+backup -- is captured once in a
    CUDA graph and replayed `sims` times) or any callable following the evaluator protocol (eager loop)."""

    def __init__(self, evaluator, sims=None, c_puct=C_PUCT, device=None, use_graph=True):
        self.model = evaluator if hasattr(evaluator, "flat_parameters") else None
        self.evaluator = network_evaluator(evaluator) if hasattr(evaluator, "predict_batch") else evaluator
        self.sims = sims
        self.c_puct = float(c_puct)
        self.device = gl._dev(device)
        self.use_graph = use_graph
        self._ws = None
        self._buf = {}
        self._graphs = {}
        self._warm = False   # one eager simulation step has run (lazy initialisation is done)

    def _workspace(self, G, max_nodes):
        L = _lib.load()
        need = L.aq_mcts_ws_bytes(G, max_nodes)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
            self._graphs.clear()  # graphs hold raw pointers into the old workspace
        return self._ws

    def _buffers(self, G):
        if G not in self._buf:
            dev = self.device
            if len(self._buf) >= 16:  # self-play shrinks G almost every ply: bound the cache; the captured graphs hold raw
                self._buf.clear()     # pointers into these buffers, so both caches are dropped together
                self._graphs.clear()
            self._buf[G] = {
                "leaf": torch.empty((G, gl.STATE_BYTES), dtype=torch.uint8, device=dev),
                "kind": torch.empty((G,), dtype=torch.int32, device=dev),
                "priors": torch.empty((G, gl.NUM_ACTIONS), dtype=torch.float32, device=dev),
                "value": torch.empty((G,), dtype=torch.float32, device=dev),
                "mask": torch.empty((G, 8), dtype=torch.int32, device=dev),
                "pawn": torch.empty((G, 8), dtype=torch.uint8, device=dev),
                "pooled": torch.empty((max(1, _lib.load().aq_leaf_eval_ws_floats(G)),), dtype=torch.float32, device=dev),  # leaf-eval workspace
            }
        return self._buf[G]

    def _step_network(self, ws, G, max_nodes, buf, flat, prep, prec, st, steps=1):
        """`steps` simulations for all games with the network evaluator: select, then `steps` leaf evaluations (three kernels each)
        joined by the fused expand + backup + next select, then the last expand + backup."""
        L, P = _lib.load(), _lib.ptr
        _lib.check(L.aq_mcts_select(P(ws), G, max_nodes, self.c_puct, P(buf["leaf"]), P(buf["kind"]), st), "aq_mcts_select")
        for i in range(steps):
            _lib.check(L.aq_leaf_eval(P(flat), P(prep), P(buf["leaf"]), G, P(buf["priors"]), P(buf["value"]), P(buf["mask"]), P(buf["pawn"]),
                                      P(buf["pooled"]), prec, st), "aq_leaf_eval")
            if i + 1 < steps:
                _lib.check(L.aq_mcts_expand_select(P(ws), G, max_nodes, P(buf["priors"]), P(buf["value"]), P(buf["mask"]), P(buf["pawn"]),
                                                   self.c_puct, P(buf["leaf"]), P(buf["kind"]), st), "aq_mcts_expand_select")
            else:
                _lib.check(L.aq_mcts_expand_backup(P(ws), G, max_nodes, P(buf["priors"]), P(buf["value"]), P(buf["mask"]),
                                                   P(buf["pawn"]), st), "aq_mcts_expand_backup")

    def _capture(self, enqueue):
        """Capture enqueue(stream pointer) into a CUDA graph.  `torch.cuda.graph` is not used on purpose: on entry it empties the
        caching allocator (the 3.5 GB workspace of a finished searcher goes back to the driver and the next one pays cudaMalloc
        again) -- 50-350 ms per capture on the boxes measured, and self-play captures once per batch size."""
        dev = self.device
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            g.capture_begin(capture_error_mode="thread_local")
            try:
                enqueue(_lib.stream_ptr(dev))
            finally:
                g.capture_end()
        torch.cuda.current_stream(dev).wait_stream(side)
        return g

    @torch.no_grad()
    def search_async(self, roots, sims=None):
        """search() without the synchronisation: returns (counts, actions, n_children, status int32[2] on the device); the caller
        passes status.tolist() to check_status() when it synchronises anyway."""
        sims = sims or self.sims or PV_EVALUATE_COUNT
        L, P = _lib.load(), _lib.ptr
        dev = self.device
        roots = roots.to(dev).contiguous()
        G_real = roots.shape[0]
        if G_real == 0:
            return (torch.empty((0, gl.MAX_LEGAL), dtype=torch.int32, device=dev), torch.empty((0, gl.MAX_LEGAL), dtype=torch.int16, device=dev),
                    torch.empty((0,), dtype=torch.int16, device=dev), torch.zeros((2,), dtype=torch.int32, device=dev))
        # Self-play and the arena drop finished games, so the number of roots shrinks by a few almost every ply.  Batches above 256
        # are padded to the next multiple of 256 with copies of the last root (searched and discarded; games are independent, so the
        # real ones are unaffected): workspaces, buffers and the captured CUDA graph of the simulation step are then reused across
        # plies instead of being rebuilt for every distinct size.
        if self.model is not None and self.use_graph and G_real > 256 and G_real % 256:
            pad = 256 - G_real % 256
            roots = torch.cat([roots, roots[-1:].expand(pad, -1)]).contiguous()
        G = roots.shape[0]
        max_nodes = 1 + sims * MAX_CHILDREN
        ws = self._workspace(G, max_nodes)
        buf = self._buffers(G)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(L.aq_mcts_reset(P(ws), P(roots), G, max_nodes, st), "aq_mcts_reset")
            if self.model is not None:
                from .pv_network_gnn import PRECISIONS
                flat = self.model.flat_parameters()
                prec = PRECISIONS[self.model.precision]
                # bf16 operand tiles, refreshed in place if the parameters changed since the last search
                prep = self.model.prepared_weights() if prec == 1 else None
                key = (G, max_nodes, flat.data_ptr(), prep.data_ptr() if prep is not None else 0, prec)
                graphs = self._graphs.get(key) if self.use_graph else None
                done = 0
                if self.use_graph:
                    if graphs is None:
                        if not self._warm:
                            # the searcher's first step runs outside capture (= simulation 1): lazy initialisation (module load,
                            # function attributes) must not happen inside one; a failure here is a real error and propagates
                            self._step_network(ws, G, max_nodes, buf, flat, prep, prec, st)
                            done, self._warm = 1, True
                            torch.cuda.synchronize(dev)
                        if len(self._graphs) >= 32:
                            self._graphs.clear()
                        self._graphs[key] = graphs = {}
                    left = sims - done
                    try:
                        # GRAPH_CHUNK simulations back to back in one graph (+ a one-step graph for the remainder, captured only when
                        # there is one): a search is a handful of graph launches instead of one per simulation -- a launch costs the
                        # host 10-100 us depending on the box, the step 75 us of GPU time at 4,096 games
                        for steps in ((GRAPH_CHUNK,) if left >= GRAPH_CHUNK else ()) + ((1,) if left % GRAPH_CHUNK else ()):
                            if steps not in graphs:
                                graphs[steps] = self._capture(lambda cst, k=steps: self._step_network(ws, G, max_nodes, buf, flat, prep, prec, cst, k))
                    except _lib.AqError:
                        raise
                    except Exception as e:  # capture unsupported: stay eager, but say so
                        import warnings
                        warnings.warn(f"CUDA-graph capture of the MCTS simulation step failed ({e!r}); running the step eagerly")
                        self.use_graph, graphs = False, None
                        self._graphs.clear()
                        torch.cuda.synchronize(dev)
                left = sims - done
                if graphs is not None:
                    for _ in range(left // GRAPH_CHUNK):
                        graphs[GRAPH_CHUNK].replay()
                    for _ in range(left % GRAPH_CHUNK):
                        graphs[1].replay()
                else:
                    for _ in range(left):
                        self._step_network(ws, G, max_nodes, buf, flat, prep, prec, st)
            else:
                for _ in range(sims):
                    _lib.check(L.aq_mcts_select(P(ws), G, max_nodes, self.c_puct, P(buf["leaf"]), P(buf["kind"]), st), "aq_mcts_select")
                    out = self.evaluator(buf["leaf"])  # terminal leaves are evaluated too and ignored by the backup
                    _lib.check(L.aq_mcts_expand_backup(P(ws), G, max_nodes, P(out["priors"]), P(out["value"]), P(out["mask"]),
                                                       P(out["pawn"]), st), "aq_mcts_expand_backup")
            counts = torch.empty((G, gl.MAX_LEGAL), dtype=torch.int32, device=dev)
            actions = torch.empty((G, gl.MAX_LEGAL), dtype=torch.int16, device=dev)
            n = torch.empty((G,), dtype=torch.int16, device=dev)
            ovf = torch.zeros((1,), dtype=torch.int32, device=dev)
            _lib.check(L.aq_mcts_root_counts(P(ws), G, max_nodes, P(counts), P(actions), P(n), P(ovf), st),
                       "aq_mcts_root_counts")
            bad = torch.isnan(buf["value"]).any().to(torch.int32).reshape(1) if self.model is not None else torch.zeros_like(ovf)
        return counts[:G_real], actions[:G_real], n[:G_real], torch.cat([ovf, bad])

    @staticmethod
    def check_status(status):
        """status: the two flags search_async returns, as host integers."""
        if status[0]:
            raise _lib.AqError("MCTS node arena overflow")
        if status[1]:
            raise _lib.AqError("the network returned NaN for a leaf: with precision 'bf16' that is how an activation beyond the fp16 range of the "
                               "aggregation operand is reported (gnn_tc2.cu); evaluate this network with precision 'fp32'")

    def search(self, roots, sims=None):
        """roots: packed uint8[G,32] (non-terminal states).  Runs `sims` simulations per game
        (pv_mcts.py:84) and returns (counts int32[G,136], actions int16[G,136], n_children int16[G]):
        visit counts of the root's children in State.legal_actions() order (pv_mcts.py:88)."""
        counts, actions, n, status = self.search_async(roots, sims)
        self.check_status(status.tolist())   # one synchronisation for both flags
        return counts, actions, n


class _SearcherCache(dict):
    """The searchers kept on a network object.  A copy or a pickle of the network starts with none: node arenas, captured graphs and
    raw pointers belong to the object they were made for."""

    def __deepcopy__(self, memo):
        return _SearcherCache()

    def __reduce__(self):
        return (_SearcherCache, ())


def searcher_for(evaluator, sims, device=None):
    """The BatchedMCTS of a network for `sims` simulations, kept on the network object: its node arenas (3.5 GB at 4,096 games x
    200 simulations), leaf buffers and captured graphs are reused by every later self-play / arena / policy call on that network
    instead of being allocated and captured again.  Evaluators that are plain callables get a fresh searcher."""
    dev = gl._dev(device)
    if not hasattr(evaluator, "flat_parameters"):
        return BatchedMCTS(evaluator, sims, device=dev)
    cache = evaluator.__dict__.setdefault("_searchers", _SearcherCache())
    key = (int(sims), str(dev))
    if key not in cache:
        if len(cache) >= 4:
            cache.clear()
        cache[key] = BatchedMCTS(evaluator, sims, device=dev)
    return cache[key]


def policy_from_counts(counts, temperature):
    """pv_mcts.py:88-95 for a batch: counts [G,136] -> float64 probabilities [G,136] (0 beyond the
    root's children)."""
    c = counts.to(torch.float64)
    if temperature == 0:  # most-visited child, first maximum (np.argmax)
        pol = torch.zeros_like(c)
        pol[torch.arange(c.shape[0], device=c.device), torch.argmax(counts, dim=1)] = 1.0
        return pol
    x = c ** (1.0 / temperature)
    return x / x.sum(dim=1, keepdim=True)


def pv_mcts_policy_batch(model, packed_roots, temperature, sims=None, device=None):
    """Batched pv_mcts_policy: -> (policy float64[G,136], actions int16[G,136], n_children int16[G])."""
    mcts = searcher_for(model, sims or PV_EVALUATE_COUNT, device)
    counts, actions, n = mcts.search(packed_roots)
    return policy_from_counts(counts, temperature), actions, n


def pv_mcts_policy(model, state, temperature, device=None):
    """Use PUCT-based Monte Carlo Tree Search to return an improved policy (distribution over legal
    actions), compared to the prior policy provided by the neural network (pv_mcts.py:20-95)."""
    rows, plies = gl.rows_from_states([state])
    packed = gl.pack_rows(rows, plies, device)
    pol, _, n = pv_mcts_policy_batch(model, packed, temperature, PV_EVALUATE_COUNT, device)
    k = int(n[0])
    out = pol[0, :k].cpu().numpy()
    return out if temperature == 0 else out.tolist()


# Action selection with Monte Carlo Tree Search
def pv_mcts_action(model, temperature=0, device=None):
    """Returns a function of the game state that selects an action based on PV-MCTS (pv_mcts.py:98-103)."""
    def pv_mcts_action(state):
        policy = pv_mcts_policy(model, deepcopy(state), temperature, device)
        return np.random.choice(state.legal_actions(), p=policy)
    return pv_mcts_action


# Boltzmann distribution
def boltzman(xs, temperature):
    """Boltzmann distribution (pv_mcts.py:106-109)"""
    xs = [x ** (1 / temperature) for x in xs]
    return [x / sum(xs) for x in xs]
