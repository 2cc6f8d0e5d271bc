"""Policy-Value network with a GNN architecture -- drop-in for reference pv_network_gnn.py.

Same module-level names (NUM_FEATURES, HIDDEN_DIM, NUM_GCN_LAYERS, POLICY_OUTPUT_SIZE,
GraphPolicyValueNetwork, create_network) and the same state_dict keys as the reference
(pv_network_gnn.py:17-80), plus ``GNNNetwork``: the BaseNetwork API (BaseNetwork.py:9-54: name,
prep_for_inference, predict, train_model, preprocess_input) with the behaviour the reference
defines in its only concrete network (pv_network_cnn.py:88-140).

All arithmetic runs in libaqgnn.so (hand-written sm_100a kernels); torch is the tensor/autograd
shell.  There is no CPU fallback: calling the network without a CUDA device raises.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import game_logic as gl
from .constants import BOARD_SIZE

# Parameters (pv_network_gnn.py:17-20)
NUM_FEATURES = 6  # Node feature size (similar to input channels in CNN)
HIDDEN_DIM = 128  # Hidden dimension for GCN layers
NUM_GCN_LAYERS = 3  # Number of GCN layers
POLICY_OUTPUT_SIZE = BOARD_SIZE ** 2 + 2 * (BOARD_SIZE - 1) ** 2  # Number of possible actions

PRECISIONS = {"fp32": 0, "bf16": 1}

# Order of the flat f32[64082] parameter buffer the kernels read (csrc/gnn_layout.cuh).  Fixed by
# NAME: nn.Module.parameters() yields a GCNConv's own `bias` before its child `lin.weight`.
FLAT_PARAM_ORDER = (
    [n for i in range(NUM_GCN_LAYERS) for n in (f"gcn_layers.{i}.lin.weight", f"gcn_layers.{i}.bias")]
    + ["policy_head.0.weight", "policy_head.0.bias", "policy_head.2.weight", "policy_head.2.bias",
       "value_head.0.weight", "value_head.0.bias", "value_head.2.weight", "value_head.2.bias"])


class _GlorotLinear(nn.Module):
    """Weight holder of GCNConv.lin (PyG Linear(bias=False, weight_initializer='glorot'))."""

    def __init__(self, n_in, n_out):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n_out, n_in))
        a = math.sqrt(6.0 / (n_in + n_out))
        nn.init.uniform_(self.weight, -a, a)


class GCNConv(nn.Module):
    """Parameter holder with the key names of torch_geometric.nn.GCNConv (lin.weight, bias)."""

    def __init__(self, n_in, n_out):
        super().__init__()
        self.lin = _GlorotLinear(n_in, n_out)
        self.bias = nn.Parameter(torch.zeros(n_out))


class _GnnFunction(torch.autograd.Function):
    """forward = aq_gnn_forward with a saved-activation workspace, backward = aq_gnn_backward."""

    @staticmethod
    def forward(ctx, flat, packed, x, open_mask, prec, *params):
        L = _lib.load()
        dev = flat.device
        B = packed.shape[0] if packed is not None else open_mask.shape[0]
        policy = torch.empty((B, POLICY_OUTPUT_SIZE), dtype=torch.float32, device=dev)
        value = torch.empty((B,), dtype=torch.float32, device=dev)
        saved = torch.empty((L.aq_gnn_saved_floats(B),), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.aq_gnn_forward(_lib.ptr(flat), _lib.ptr(packed), _lib.ptr(x), _lib.ptr(open_mask), B,
                                        _lib.ptr(policy), _lib.ptr(value), _lib.ptr(saved), prec, _lib.stream_ptr(dev)),
                       "aq_gnn_forward")
        ctx.save_for_backward(flat, saved)
        ctx.B = B
        ctx.prec = prec
        ctx.shapes = [p.shape for p in params]
        return policy, value.unsqueeze(1)

    @staticmethod
    def backward(ctx, dpolicy, dvalue):
        L = _lib.load()
        flat, saved = ctx.saved_tensors
        dev, B = flat.device, ctx.B
        dpolicy = dpolicy.contiguous().float()
        dvalue = dvalue.reshape(B).contiguous().float()
        grads = torch.empty_like(flat)
        ws = torch.empty((L.aq_gnn_backward_ws_floats(B),), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.aq_gnn_backward(_lib.ptr(flat), _lib.ptr(saved), _lib.ptr(dpolicy), _lib.ptr(dvalue), B,
                                         _lib.ptr(grads), _lib.ptr(ws), ctx.prec, _lib.stream_ptr(dev)), "aq_gnn_backward")
        out, off = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp))
            out.append(grads[off:off + n].view(shp))
            off += n
        return (None, None, None, None, None, *out)


# Graph-based Policy-Value Network
class GraphPolicyValueNetwork(nn.Module):
    def __init__(self, num_features, hidden_dim, num_gcn_layers, policy_output_size):
        super(GraphPolicyValueNetwork, self).__init__()
        if (num_features, hidden_dim, num_gcn_layers, policy_output_size) != (
                NUM_FEATURES, HIDDEN_DIM, NUM_GCN_LAYERS, POLICY_OUTPUT_SIZE):
            raise ValueError("libaqgnn.so is compiled for num_features=6, hidden_dim=128, num_gcn_layers=3, "
                             "policy_output_size=209 (the reference's constants, pv_network_gnn.py:17-20)")
        self.num_features = num_features
        self.hidden_dim = hidden_dim
        self.num_gcn_layers = num_gcn_layers
        self.policy_output_size = policy_output_size

        # GCN layers
        self.gcn_layers = nn.ModuleList()
        self.gcn_layers.append(GCNConv(num_features, hidden_dim))
        for _ in range(num_gcn_layers - 1):
            self.gcn_layers.append(GCNConv(hidden_dim, hidden_dim))

        # Policy head (containers for the parameters; Softmax/Tanh are applied inside the kernel)
        self.policy_head = nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim // 2),
            nn.ReLU(),
            nn.Linear(hidden_dim // 2, policy_output_size),
            nn.Softmax(dim=1)
        )

        # Value head
        self.value_head = nn.Sequential(
            nn.Linear(hidden_dim, hidden_dim // 2),
            nn.ReLU(),
            nn.Linear(hidden_dim // 2, 1),
            nn.Tanh()
        )
        self.precision = "fp32"  # inference arithmetic: "fp32" (FFMA) or "bf16" (tcgen05 tensor cores)
        self.train_precision = "fp32"  # forward+backward under autograd: "fp32", or "bf16" (tcgen05 trunk, fp32 accumulate)
        self._flat = None
        self._prepared = None      # bf16 operand tiles of the tensor-core inference kernels (aq_prepare_inference)
        self._prepared_key = None
        self._weights_epoch = 0    # bumped by writers that update the flat buffer through raw pointers (FlatTrainer)

    def ordered_parameters(self):
        # the Parameter objects keep their identity across .to() / load_state_dict (only .data is replaced), so the module
        # traversal is done once; this is on the path of every predict / search call
        cached = self.__dict__.get("_ordered_params")
        if cached is None:
            named = dict(self.named_parameters())
            cached = self.__dict__["_ordered_params"] = [named[n] for n in FLAT_PARAM_ORDER]
        return cached

    # ---- flat parameter buffer: every parameter is a view into one f32[64082] tensor ------------
    def flat_parameters(self):
        """The flat parameter buffer the kernels read (state_dict order).  Parameters are re-pointed
        into it lazily, e.g. after .to(device)."""
        params = self.ordered_parameters()
        dev = params[0].device
        flat = self._flat
        ok = flat is not None and flat.device == dev
        if ok:
            off = 0
            for p in params:
                ok = ok and p.data_ptr() == flat.data_ptr() + 4 * off and p.dtype == torch.float32
                off += p.numel()
        if not ok:
            total = sum(p.numel() for p in params)
            assert total == _lib.load().aq_param_count(), "parameter count does not match libaqgnn.so"
            flat = torch.empty((total,), dtype=torch.float32, device=dev)
            off = 0
            for p in params:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                off += n
            self._flat = flat
        return flat

    def mark_weights_changed(self):
        """Tell the network its flat buffer was modified behind torch's back (raw-pointer Adam step)."""
        self._weights_epoch += 1

    def prepared_weights(self):
        """Device buffer with the inference weights in the layout the tensor-core kernels keep in shared
        memory, rebuilt (one small kernel) whenever the parameters changed.  The counterpart of the
        reference's one-time inference preparation (BaseNetwork.py:22-32)."""
        flat = self.flat_parameters()
        _lib.require_cuda(flat, "model parameters")
        key = (flat.data_ptr(), flat._version, self._weights_epoch, tuple(p._version for p in self.ordered_parameters()))
        if self._prepared is None or self._prepared.device != flat.device or key != self._prepared_key:
            L = _lib.load()
            if self._prepared is None or self._prepared.device != flat.device:
                self._prepared = torch.empty((L.aq_prepared_bytes(),), dtype=torch.uint8, device=flat.device)
            with torch.cuda.device(flat.device):
                _lib.check(L.aq_prepare_inference(_lib.ptr(flat), _lib.ptr(self._prepared), _lib.stream_ptr(flat.device)),
                           "aq_prepare_inference")
            self._prepared_key = key
        return self._prepared

    def _prepare(self, x, edge_index, batch):
        """-> (packed, x, open_mask): either packed states or explicit graph inputs on the device."""
        flat = self.flat_parameters()
        _lib.require_cuda(flat, "model parameters")
        dev = flat.device
        if edge_index is None:
            # states: packed uint8[B,32], or State.to_array() rows [B,68] of any numeric dtype
            s = torch.as_tensor(x)
            if s.dim() != 2 or s.shape[1] not in (gl.STATE_BYTES, 68):
                raise ValueError("expected packed states [B,32] or row68 states [B,68]")
            if s.shape[1] == 68:
                packed = gl.pack_rows(s.to(torch.uint8), None, dev)
            else:
                packed = s.to(device=dev, dtype=torch.uint8).contiguous()
            return flat, packed, None, None
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 2 or x.shape[1] != NUM_FEATURES or x.shape[0] % gl.NUM_SQUARES != 0:
            raise ValueError("x must be [B*81, 6]")
        B = x.shape[0] // gl.NUM_SQUARES
        if batch is not None:
            expect = torch.arange(B, device=batch.device).repeat_interleave(gl.NUM_SQUARES)
            if batch.shape[0] != x.shape[0] or not torch.equal(batch.to(torch.int64), expect):
                raise ValueError("batch must assign 81 consecutive nodes to each graph")
        open_mask = gl.open_mask_from_edge_index(edge_index.to(dev), B)
        return flat, None, x, open_mask

    def forward(self, x, edge_index=None, batch=None):
        """forward(x, edge_index, batch) as in the reference (pv_network_gnn.py:53-64), or
        forward(states) with packed / row68 states (graph built inside the kernel).
        Returns (policy [B,209] softmax probabilities, value [B,1])."""
        flat, packed, xx, open_mask = self._prepare(x, edge_index, batch)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            prec = PRECISIONS[self.train_precision] if packed is not None else 0
            return _GnnFunction.apply(flat, packed, xx, open_mask, prec, *self.ordered_parameters())
        L = _lib.load()
        dev = flat.device
        B = packed.shape[0] if packed is not None else open_mask.shape[0]
        policy = torch.empty((B, POLICY_OUTPUT_SIZE), dtype=torch.float32, device=dev)
        value = torch.empty((B,), dtype=torch.float32, device=dev)
        prec = PRECISIONS[self.precision] if packed is not None else 0
        with torch.cuda.device(dev):
            _lib.check(L.aq_gnn_forward(_lib.ptr(flat), _lib.ptr(packed), _lib.ptr(xx), _lib.ptr(open_mask), B,
                                        _lib.ptr(policy), _lib.ptr(value), None, prec, _lib.stream_ptr(dev)),
                       "aq_gnn_forward")
        return policy, value.unsqueeze(1)


class GNNNetwork(GraphPolicyValueNetwork):
    """BaseNetwork API (BaseNetwork.py:9-54) for the GNN; 0-argument constructor like CNNNetwork."""

    def __init__(self):
        super().__init__(NUM_FEATURES, HIDDEN_DIM, NUM_GCN_LAYERS, POLICY_OUTPUT_SIZE)
        self._name = 'GNN'
        self.optimised_model = None  # BaseNetwork.py:13; the CUDA kernels are the optimised model

    @property
    def name(self):
        return self._name

    def prep_for_inference(self, model_path):
        """Loads state_dict of parameters located at model_path and prepares the model for
        inference (BaseNetwork.py:21-32; TensorRT compilation is replaced by libaqgnn)."""
        if not torch.cuda.is_available():
            raise _lib.AqError("prep_for_inference needs a CUDA device (no CPU fallback)")
        self.load_state_dict(torch.load(model_path, map_location='cuda'))
        self.eval()
        self.to('cuda')
        self.flat_parameters()
        self.prepared_weights()
        self.optimised_model = self

    def preprocess_input(self, game_state_arrays):
        """list of State.to_array() triples -> uint8[M,68] rows, the form ``forward`` accepts
        (BaseNetwork.py:49-54; the CNN's version builds 6 planes, pv_network_cnn.py:88-114 -- here
        the planes and the graph are built on the GPU from these rows)."""
        return gl.rows_from_arrays(game_state_arrays)

    @torch.no_grad()
    def predict_batch(self, states, plies=None):
        """Batched predict for B leaves.  states: list of State objects, row68 array/tensor [B,68], or
        packed CUDA tensor [B,32].  Returns dict(priors f32[B,209] legal-masked and renormalised,
        value f32[B], mask int32[B,8], pawn uint8[B,8]) on the device."""
        flat = self.flat_parameters()
        _lib.require_cuda(flat, "model parameters")
        dev = flat.device
        if isinstance(states, (list, tuple)):
            rows, plies = gl.rows_from_states(states)
            packed = gl.pack_rows(rows, plies, dev)
        else:
            s = torch.as_tensor(states)
            packed = gl.pack_rows(s, plies, dev) if s.shape[1] == 68 else s.to(dev).contiguous()
        B = packed.shape[0]
        L = _lib.load()
        priors = torch.empty((B, POLICY_OUTPUT_SIZE), dtype=torch.float32, device=dev)
        value = torch.empty((B,), dtype=torch.float32, device=dev)
        mask = torch.empty((B, 8), dtype=torch.int32, device=dev)
        pawn = torch.empty((B, 8), dtype=torch.uint8, device=dev)
        ws = torch.empty((max(1, L.aq_leaf_eval_ws_floats(B)),), dtype=torch.float32, device=dev)
        prep = self.prepared_weights() if self.precision == "bf16" else None
        with torch.cuda.device(dev):
            _lib.check(L.aq_leaf_eval(_lib.ptr(flat), _lib.ptr(prep), _lib.ptr(packed), B, _lib.ptr(priors), _lib.ptr(value),
                                      _lib.ptr(mask), _lib.ptr(pawn), _lib.ptr(ws), PRECISIONS[self.precision],
                                      _lib.stream_ptr(dev)), "aq_leaf_eval")
        return {"priors": priors, "value": value, "mask": mask, "pawn": pawn, "packed": packed}

    def predict(self, state, device=None):
        """Predict the policy and value for a game state given a State object.
        :returns: policy, value, where policy is a normalised PMF over all legal actions (1D numpy
        array in state.legal_actions() order) and value is a float between -1 and 1
        (BaseNetwork.py:36-40; pv_network_cnn.py:117-137)."""
        out = self.predict_batch([state])
        actions, n = gl.legal_actions_batch(out["packed"], out["mask"], out["pawn"])
        idx = actions[0, : int(n[0])].to(torch.int64)
        policy = out["priors"][0][idx].cpu().numpy()
        return policy, out["value"][0].item()

    def train_model(self, data_loader, optimizer, loss_fn=None, device='cuda', num_epochs=10):
        """Performs model training (BaseNetwork.py:42-45).  Batches are (states, policy_target,
        value_target) as produced by train_network.py:42-49; the default loss is the reference's
        CrossEntropyLoss-on-softmax + MSELoss (train_network.py:54-55,85-89)."""
        if loss_fn is None:
            ce, mse = nn.CrossEntropyLoss(), nn.MSELoss()
            loss_fn = lambda pp, vp, pt, vt: ce(pp, pt) + mse(vp.squeeze(), vt)  # noqa: E731
        self.to(device)
        history = []
        for _ in range(num_epochs):
            self.train()
            total = 0.0
            for state, policy_target, value_target in data_loader:
                policy_pred, value_pred = self(state.to(device))
                loss = loss_fn(policy_pred, value_pred, policy_target.to(device), value_target.to(device))
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
                total += loss.item()
            history.append(total)
        return history


class HostLeafEvaluator:
    """Batched ``predict`` for callers whose states and results live in HOST memory (a Python tree search such as the
    reference's pv_mcts.py Node tree, many games at once): one call moves B packed states to the GPU, evaluates them and
    brings back exactly what ``BaseNetwork.predict`` returns per state -- the probabilities of the legal actions only, in
    ``state.legal_actions()`` order (BaseNetwork.py:36-40) -- as a ragged array, plus the values.

    Pinned host buffers, the device workspace and the worker streams are allocated once for ``max_batch``.
        ev = HostLeafEvaluator(net, 16384)
        ev.states[:B] = game_logic.pack_rows_host(rows, plies)      # fill the pinned input
        out = ev.evaluate(B)   # dict of numpy views: priors [total], offsets [B+1], value [B] (+ mask [B,8], pawn [B,8])
        p_b = out["priors"][out["offsets"][b]:out["offsets"][b + 1]]   # == predict(state_b)[0]

    wire:      "f32" -- the ragged priors are the float32 bits of the dense priors (exact);
               "f16" -- IEEE half on the wire and in the returned array (relative error <= 2^-12, well inside the stated
                        bf16-path tolerance; halves the device -> host bytes, which bound the multi-GPU host path).
    with_mask: also return the 256-bit legal masks and ordered pawn lists (the action ids of the ragged entries without a host
               ``legal_actions()`` call); predict() itself returns only (policy, value).
    dense:     the [B,209] matrix instead of the ragged array (aq_leaf_eval_host).
    """

    def __init__(self, net, max_batch, dense=False, wire="f32", with_mask=True):
        import ctypes
        flat = net.flat_parameters()
        _lib.require_cuda(flat, "model parameters")
        if wire not in ("f32", "f16") or (dense and wire != "f32"):
            raise ValueError("wire must be 'f32' or 'f16' (ragged flavour only)")
        self.net, self.dev, self.max_batch, self.dense = net, flat.device, int(max_batch), bool(dense)
        self.wire, self.with_mask = wire, bool(with_mask) or bool(dense)
        L = self.L = _lib.load()
        B = self.max_batch
        self.states = torch.empty((B, gl.STATE_BYTES), dtype=torch.uint8).pin_memory()
        self.value = torch.empty((B,), dtype=torch.float32).pin_memory()
        self.mask = torch.empty((B, 8), dtype=torch.int32).pin_memory() if self.with_mask else None
        self.pawn = torch.empty((B, 8), dtype=torch.uint8).pin_memory() if self.with_mask else None
        if dense:
            self.priors = torch.empty((B, POLICY_OUTPUT_SIZE), dtype=torch.float32).pin_memory()
            self.offsets = None
            nbytes = L.aq_leaf_eval_host_ws_bytes(B)
        else:
            # one pinned block laid out like the device's result block (offsets | value | ragged priors): one D2H copy per batch
            lay = (ctypes.c_int64 * 2)()
            _lib.check(L.aq_leaf_eval_host_compact_layout(B, lay), "aq_leaf_eval_host_compact_layout")
            elem = 2 if wire == "f16" else 4
            self._block = torch.empty((lay[1] + B * gl.MAX_LEGAL * elem,), dtype=torch.uint8).pin_memory()
            self.offsets = self._block[: (B + 1) * 4].view(torch.int32)
            self.value = self._block[lay[0]: lay[0] + B * 4].view(torch.float32)
            self.priors = self._block[lay[1]:].view(torch.float16 if wire == "f16" else torch.float32)
            nbytes = L.aq_leaf_eval_host_compact_ws_bytes(B)
        self.ws = torch.empty((max(1, nbytes),), dtype=torch.uint8, device=self.dev)
        self._ctx = ctypes.c_void_p()
        self._inflight = None
        _lib.check(L.aq_host_ctx_create(ctypes.byref(self._ctx)), "aq_host_ctx_create")
        self.refresh_weights()

    def refresh_weights(self):
        """Re-reads the network's parameters (flat buffer, bf16 operand tiles, precision).  Like the reference's
        prep_for_inference (BaseNetwork.py:21-32) the evaluator works on the weights as they were at this call; call it again
        after a training step or load_state_dict."""
        net = self.net
        self._flat = net.flat_parameters()
        self._prec = PRECISIONS[net.precision]
        self._prep = net.prepared_weights() if self._prec == 1 else None
        P = _lib.ptr
        self._fixed = (P(self.priors), P(self.offsets) if self.offsets is not None else None, P(self.value), P(self.mask), P(self.pawn),
                       P(self.ws), P(self._flat), P(self._prep), P(self.states))

    def close(self):
        if self._ctx:
            self.L.aq_host_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _args(self, B, states):
        src = self.states if states is None else states
        if src.is_cuda or src.dtype != torch.uint8 or not src.is_contiguous():
            raise ValueError("states must be a contiguous host uint8[B,32] tensor")
        B = (self.max_batch if states is None else states.shape[0]) if B is None else int(B)
        if not 0 <= B <= min(self.max_batch, src.shape[0]):
            raise ValueError("batch larger than max_batch")
        return B, (self._fixed[8] if states is None else _lib.ptr(src))

    def evaluate(self, B=None, states=None):
        """Evaluates self.states[:B] (or ``states``, the caller's own host uint8[B,32] tensor -- pinned memory makes the copy
        asynchronous); synchronous.  Returns numpy views into the pinned result buffers."""
        if self.dense:
            B, p_states = self._args(B, states)
            p_pri, _, p_val, p_msk, p_pwn, p_ws, p_flat, p_prep, _ = self._fixed
            with torch.cuda.device(self.dev):
                _lib.check(self.L.aq_leaf_eval_host(p_flat, p_prep, p_states, B, p_pri, p_val, p_msk, p_pwn, p_ws, self._prec, self._ctx,
                                                    _lib.stream_ptr(self.dev)), "aq_leaf_eval_host")
            return {"priors": self.priors[:B].numpy(), "value": self.value[:B].numpy(), "mask": self.mask[:B].numpy(),
                    "pawn": self.pawn[:B].numpy()}
        self.submit(B, states)
        return self.wait()

    def submit(self, B=None, states=None):
        """First half of ``evaluate`` (ragged flavour only): enqueue the copies and kernels of this batch and return at once.
        A host that keeps several evaluators busy round-robin -- ``a.submit(); b.submit(); c.submit(); a.wait(); a.submit();
        b.wait(); ...`` -- hides each batch's transfers and wake-up latency behind the others' kernels."""
        if self.dense:
            raise ValueError("submit/wait exist for the ragged (predict-shaped) flavour")
        B, p_states = self._args(B, states)
        p_pri, p_off, p_val, p_msk, p_pwn, p_ws, p_flat, p_prep, _ = self._fixed
        with torch.cuda.device(self.dev):
            _lib.check(self.L.aq_leaf_eval_host_compact_submit(p_flat, p_prep, p_states, B, p_pri, self.priors.numel(), p_off, p_val, p_msk,
                                                               p_pwn, p_ws, self._prec, 1 if self.wire == "f16" else 0, self._ctx,
                                                               _lib.stream_ptr(self.dev)),
                       "aq_leaf_eval_host_compact_submit")
        self._inflight = B

    def wait(self):
        """Second half: returns when the results of the submitted batch are on the host (numpy views into the pinned buffers)."""
        B = self._inflight
        if B is None:
            raise ValueError("wait() without submit()")
        self._inflight = None
        with torch.cuda.device(self.dev):
            _lib.check(self.L.aq_leaf_eval_host_compact_wait(self._ctx), "aq_leaf_eval_host_compact_wait")
        off = self.offsets[:B + 1].numpy()
        out = {"priors": self.priors[:int(off[B])].numpy(), "offsets": off, "value": self.value[:B].numpy()}
        if self.with_mask:
            out["mask"], out["pawn"] = self.mask[:B].numpy(), self.pawn[:B].numpy()
        return out

    def stats(self):
        """(batches whose ragged copy had to be completed by a second copy, current estimate of legal actions per board)."""
        import ctypes
        out = (ctypes.c_int64 * 2)()
        _lib.check(self.L.aq_host_ctx_stats(self._ctx, out), "aq_host_ctx_stats")
        return int(out[0]), out[1] / 1024.0

    def d2h_bytes(self, out):
        """Bytes of RESULTS that crossed PCIe device -> host for this batch (for bench.py's e2e accounting; the ragged copy is sized
        by an estimate, so up to ~3 % more than this may actually have moved)."""
        B = out["value"].shape[0]
        extra = (32 + 8) if self.with_mask else 0
        if self.dense:
            return B * (POLICY_OUTPUT_SIZE * 4 + 4 + extra)
        return int(out["offsets"][B]) * (2 if self.wire == "f16" else 4) + (B + 1) * 4 + B * (4 + extra)


# Function to create the dual network
def create_network(model_path='model/best.pth'):
    # Do nothing if the model is already created (pv_network_gnn.py:68-80)
    if os.path.exists(model_path):
        return

    # Initialize the model
    model = GraphPolicyValueNetwork(NUM_FEATURES, HIDDEN_DIM, NUM_GCN_LAYERS, POLICY_OUTPUT_SIZE)

    # Save the model
    os.makedirs(os.path.dirname(model_path) or '.', exist_ok=True)
    torch.save(model.state_dict(), model_path)


# Running the function
if __name__ == '__main__':
    create_network()
