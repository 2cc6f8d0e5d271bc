"""Network training -- drop-in for reference train_network.py (train_network.py:14-107): newest
./data/*.history, batch 128 shuffled, loss = CrossEntropyLoss(softmax output, target) + MSELoss,
Adam(lr=1e-3) with the 1.0 / 0.5 / 0.25 LambdaLR schedule, 100 epochs, weights saved to latest.pth.

The step runs entirely in libaqgnn.so on flat buffers (forward with saved activations, fused loss
gradient, atomic-free backward, Adam).  Data parallel: every rank computes its slice of each global
batch, gradients are summed with ONE all-reduce of the flat 64,082-float buffer (NCCL over NVLink on
GPUs), and every rank applies the identical Adam update."""
import os
import pickle
from pathlib import Path

import numpy as np
import torch

from . import _lib
from . import game_logic as gl
from .constants import PV_NETWORK_PATH
from .pv_network_gnn import GNNNetwork, POLICY_OUTPUT_SIZE

NUM_EPOCH = 100   # train_network.py:14
BATCH_SIZE = 128  # train_network.py:15


def load_data(directory='./data'):
    # Load the latest training data (train_network.py:18-23)
    history_path = sorted(Path(directory).glob('*.history'))[-1]
    with history_path.open(mode='rb') as f:
        return pickle.load(f)


def lr_lambda(epoch):
    """train_network.py:59-65"""
    if epoch >= 80:
        return 0.25
    elif epoch >= 50:
        return 0.5
    return 1.0


def shard_bounds(n, rank, world_size):
    """Contiguous split of a batch of n samples over the ranks (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatTrainer:
    """One model replica + Adam state on flat buffers; step() = forward/loss/backward/all-reduce/Adam."""

    def __init__(self, model, lr=0.001, betas=(0.9, 0.999), eps=1e-8, rank=0, world_size=1, precision=None):
        from .pv_network_gnn import PRECISIONS
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.prec = PRECISIONS[precision if precision is not None else model.train_precision]
        self.rank, self.world = rank, world_size
        self.flat = model.flat_parameters()
        _lib.require_cuda(self.flat, "model parameters")
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.grads = torch.zeros_like(self.flat)
        self.step_count = 0
        self.loss = torch.zeros(2, device=self.flat.device)

    def step(self, packed, policy_target, value_target, global_batch, lr_scale=1.0):
        """packed uint8[b,32], policy_target f32[b,209], value_target f32[b]: this rank's slice of a
        global batch of `global_batch` samples (b may be 0).  Returns the loss tensor [policy, value]
        contribution of this rank (already divided by the global batch size)."""
        L, P = _lib.load(), _lib.ptr
        flat = self.model.flat_parameters()
        if flat.data_ptr() != self.flat.data_ptr() or flat.device != self.flat.device:
            # the module was moved (.to) or its storage replaced (load_state_dict(assign=True)) since the last step: follow it, keep
            # the Adam moments (on the new device), so that the update does not go into an orphaned buffer
            self.flat = flat
            self.exp_avg, self.exp_avg_sq = self.exp_avg.to(flat.device), self.exp_avg_sq.to(flat.device)
            self.grads, self.loss = torch.zeros_like(flat), torch.zeros(2, device=flat.device)
        dev = self.flat.device
        b = packed.shape[0]
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            if b > 0:
                policy = torch.empty((b, POLICY_OUTPUT_SIZE), dtype=torch.float32, device=dev)
                value = torch.empty((b,), dtype=torch.float32, device=dev)
                saved = torch.empty((L.aq_gnn_saved_floats(b),), dtype=torch.float32, device=dev)
                ws = torch.empty((L.aq_gnn_backward_ws_floats(b),), dtype=torch.float32, device=dev)
                dp, dv = torch.empty_like(policy), torch.empty_like(value)
                _lib.check(L.aq_gnn_forward(P(self.flat), P(packed), None, None, b, P(policy), P(value), P(saved), self.prec, st), "aq_gnn_forward")
                _lib.check(L.aq_loss_grad(P(policy), P(value), P(policy_target), P(value_target), b, global_batch, P(self.loss),
                                          P(dp), P(dv), st), "aq_loss_grad")
                _lib.check(L.aq_gnn_backward(P(self.flat), P(saved), P(dp), P(dv), b, P(self.grads), P(ws), self.prec, st), "aq_gnn_backward")
            else:
                self.grads.zero_()
                self.loss.zero_()
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(self.grads)  # one 256 KB all-reduce per step; losses divide by the global batch already
            self.step_count += 1
            _lib.check(L.aq_adam_step(P(self.flat), P(self.grads), P(self.exp_avg), P(self.exp_avg_sq), self.flat.numel(),
                                      self.step_count, self.lr * lr_scale, self.betas[0], self.betas[1], self.eps, 1.0, st),
                       "aq_adam_step")
        if hasattr(self.model, "mark_weights_changed"):
            self.model.mark_weights_changed()  # the flat buffer was written through a raw pointer
        return self.loss


def train_on_history(model, history, num_epochs=NUM_EPOCH, batch_size=BATCH_SIZE, rank=0, world_size=1, seed=0,
                     verbose=True):
    """The training loop of train_network.py:69-104 on an in-memory history."""
    dev = model.flat_parameters().device
    s, p, v = zip(*history)
    rows = model.preprocess_input(s)                                   # train_network.py:37
    packed = gl.pack_rows(rows, None, dev)
    p = torch.tensor(np.array(p), dtype=torch.float32, device=dev)     # policy targets
    v = torch.tensor(np.array(v), dtype=torch.float32, device=dev)     # value targets
    return train_on_buffer(model, packed, p, v, num_epochs, batch_size, rank, world_size, seed, verbose)


def train_on_buffer(model, packed, p, v, num_epochs=NUM_EPOCH, batch_size=BATCH_SIZE, rank=0, world_size=1, seed=0, verbose=True):
    """The same loop on a self-play record that is already on the device (self_play.play_batch_device): packed states
    uint8[M,32], policy targets f32[M,209], value targets f32[M]."""
    dev = model.flat_parameters().device
    packed, p, v = packed.to(dev), p.to(device=dev, dtype=torch.float32), v.to(device=dev, dtype=torch.float32)
    M = packed.shape[0]
    trainer = FlatTrainer(model, lr=0.001, rank=rank, world_size=world_size)
    gen = torch.Generator(device='cpu')
    gen.manual_seed(seed)                                              # same shuffle on every rank
    losses = []
    model.train()
    for epoch in range(num_epochs):
        perm = torch.randperm(M, generator=gen).to(dev)                # DataLoader(shuffle=True)
        ep = torch.zeros(2, device=dev)
        for start in range(0, M, batch_size):
            idx = perm[start:start + batch_size]
            lo, hi = shard_bounds(idx.numel(), rank, world_size)
            mine = idx[lo:hi]
            ep += trainer.step(packed[mine].contiguous(), p[mine].contiguous(), v[mine].contiguous(), idx.numel(),
                               lr_lambda(epoch))
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(ep)
        losses.append(ep.tolist())
        if verbose and rank == 0:
            print(f"\rEpoch {epoch + 1}/{num_epochs} | Policy Loss: {losses[-1][0]:.4f} | Value Loss: {losses[-1][1]:.4f}", end='')
    if verbose and rank == 0:
        print('')
    return losses


def train_network(data_dir='./data', model_dir=None, num_epochs=NUM_EPOCH, rank=0, world_size=1):
    model_dir = PV_NETWORK_PATH if model_dir is None else model_dir
    # Load the model (train_network.py:28-30)
    model = GNNNetwork()
    model.load_state_dict(torch.load(os.path.join(model_dir, 'best.pth'), map_location='cuda'))
    model = model.to('cuda')
    history = load_data(data_dir)
    losses = train_on_history(model, history, num_epochs, BATCH_SIZE, rank, world_size)
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(model_dir, 'latest.pth'))  # train_network.py:107
    return losses


if __name__ == '__main__':
    train_network()
