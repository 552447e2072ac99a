"""Whole-step CUDA graphs for the launch-bound configurations.

At the reference's own batch size (256, config.py:31) a P-Companion training step is ~60 kernels of a few microseconds
each: the step time is host launch overhead, not GPU work.  ``GraphedTrainStep`` captures forward + loss + backward +
optimiser step once (PyTorch's whole-network capture recipe: static input buffers, warm-up on a side stream, a
capturable optimiser) and replays it with one launch.  Every C-ABI kernel launches on the capturing stream and
allocates through torch's caching allocator, so the capture needs nothing special from the native side; the dropout mask
of the type-transition layer takes its per-replay seed from a device counter that the graph itself increments.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, example_batch); loss = step(batch)

    ``optimizer`` must be capturable (e.g. ``torch.optim.Adam(..., capturable=True)``).  ``batch`` tensors are copied into
    the static buffers (host tensors: an H2D copy on the current stream); the returned loss is the graph's static
    output tensor (read it before the next call).  ``loss_fn(model, batch) -> scalar`` defaults to the P-Companion joint
    loss ``model.compute_loss(batch, model(batch))``."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, example_batch: Dict[str, torch.Tensor],
                 loss_fn: Callable = None, warmup: int = 3):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep: the model must live on a CUDA device (no CPU fallback)")
        self.model, self.optimizer = model, optimizer
        self.loss_fn = loss_fn or (lambda m, b: m.compute_loss(b, m(b)))
        self.static = {k: v.to(dev).clone() for k, v in example_batch.items() if torch.is_tensor(v)}
        self._counters = []
        for mod in model.modules():
            if hasattr(mod, "_run") and hasattr(mod, "dropout"):       # ComplementaryTypeTransition
                mod._replay_counter = torch.zeros(1, dtype=torch.int64, device=dev)
                self._counters.append(mod)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                optimizer.zero_grad(set_to_none=True)
                self.loss_fn(model, self.static).backward()
                optimizer.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self.loss_fn(model, self.static)
            self.loss.backward()
            optimizer.step()

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        for k, buf in self.static.items():
            src = batch[k]
            if src is not buf:
                buf.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss

    def release(self) -> None:
        for mod in self._counters:
            mod._replay_counter = None
