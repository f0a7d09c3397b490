"""CUDA-graph replay of continuation chunks for small-chunk streaming.

A continuation chunk of a few frames is launch bound: ~146 kernel launches per forward cost ~2.3 ms
eager at 196 tokens while the kernels themselves need ~1.3 ms (tools/graph_probe.py).  The library
only enqueues on the caller's stream and allocates nothing, so a whole forward captures as is; this
runner adds what a stream needs around it: replay-stable input / state buffers, the state fed back
inside the graph, and the temporal position rows of the current offset (reference
videomamba.py:655-675, computed eagerly per chunk so the interpolation rule is untouched) written
into a fixed buffer before each replay.

Semantics are those of ``model(x, ssm_state=state, temporal_pos_offset=offset)`` with the state of
the previous call (streaming contract 1.0.0).  The first chunk carries the CLS token, so its shapes
differ: it has a graph of its own (zero state in, captured on its first use; an explicit initial state
runs eagerly), every later chunk replays the continuation graph.  Results are bit-identical to the
eager calls (same kernels, same order).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

LayerState = Tuple[Tensor, Tensor]


class GraphedStream:
    """One batch of concurrent streams advancing ``chunk_frames`` frames per call."""

    def __init__(self, model, warmup: int = 2):
        if getattr(model, "add_pool_norm", True) and getattr(model, "pool_type", "avg") != "avg":
            raise ValueError("continuation chunks carry no CLS token: build the model with "
                             "pool_type='avg' (or add_pool_norm=False) for graphed streaming.")
        self.model = model
        self.warmup = int(warmup)
        self.offset = 0                       # temporal tokens consumed so far
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._x: Optional[Tensor] = None
        self._rows: Optional[Tensor] = None
        self._state: Optional[List[LayerState]] = None
        self._out = None
        # the graph bakes in the pointers of every mixer's kernel operands: keep those objects alive
        # and recapture when any of them was rebuilt (parameter update, refresh_weights())
        self._captured_keys: Optional[list] = None
        self._captured_weights: Optional[list] = None
        # first-chunk graph (offset 0, CLS included, zero state in)
        self._first_graph: Optional[torch.cuda.CUDAGraph] = None
        self._first_x: Optional[Tensor] = None
        self._first_out = None
        self._first_keys: Optional[list] = None
        self._first_weights: Optional[list] = None
        self._zero_state: Optional[List[LayerState]] = None

    # ---- public -------------------------------------------------------------------------
    @property
    def state(self) -> Optional[List[LayerState]]:
        """The carried ``[(conv_state, ssm_state)] * depth`` (live buffers; clone to keep)."""
        return self._state

    def first(self, x: Tensor, state=None):
        """First chunk of the streams (offset 0, CLS included).  Returns what
        ``model(x, ssm_state=state, temporal_pos_offset=0)`` returns, minus the state.  With the default
        zero state the chunk is one graph launch (the returned tensors are then the graph's output
        buffers, overwritten by the next ``first``); an explicit ``state`` runs eagerly."""
        m = self.model
        self.offset = m._validate_temporal_length(x.shape[2])
        if state is not None:
            with torch.no_grad():
                out = m(x, ssm_state=state, temporal_pos_offset=0)
            self._adopt_state(out[-1])
            return out[:-1] if len(out) > 2 else out[0]
        if self._first_graph is None or self._first_x.shape != x.shape or self._first_x.dtype != x.dtype \
                or self._first_keys != self._mixer_keys():
            self._capture_first(x)
        else:
            self._first_x.copy_(x)
            self._first_graph.replay()
        out = self._first_out
        return out[:-1] if len(out) > 2 else out[0]

    def step(self, x: Tensor):
        """Next chunk: one graph launch.  The returned tensors are the graph's output buffers and
        are overwritten by the next ``step`` (clone to keep)."""
        if self._state is None:
            raise RuntimeError("call first() before step().")
        m = self.model
        t_tokens = m._validate_temporal_length(x.shape[2])
        rows = self._temporal_rows(t_tokens, x)
        if self._graph is None or self._x.shape != x.shape or self._x.dtype != x.dtype \
                or self._captured_keys != self._mixer_keys():
            self._capture(x, rows)
        else:
            self._x.copy_(x)
            self._rows.copy_(rows)
            self._graph.replay()
        self.offset += t_tokens
        out = self._out
        return out[:-1] if len(out) > 2 else out[0]

    # ---- internals ----------------------------------------------------------------------
    def _adopt_state(self, new_state) -> None:
        """Copy a freshly returned state into the carried buffers (allocating them on first use; the
        continuation graph keeps reading the SAME buffers)."""
        new = self._as_list(new_state)
        if self._state is None or any(c.shape != nc.shape or c.dtype != nc.dtype or s.dtype != ns.dtype
                                      for (c, s), (nc, ns) in zip(self._state, new)):
            self._state = [(c.clone(), s.clone()) for c, s in new]
            self._graph = None
        else:
            for (c, s), (nc, ns) in zip(self._state, new):
                c.copy_(nc)
                s.copy_(ns)

    def _run_first(self):
        out = self.model(self._first_x, ssm_state=self._zero_state, temporal_pos_offset=0)
        self._adopt_state(out[-1])
        return out

    def _capture_first(self, x: Tensor) -> None:
        m = self.model
        p = next(m.parameters())
        self._first_x = x.clone()
        self._zero_state = m.allocate_state(x.shape[0], dtype=p.dtype, device=x.device)   # never written
        with torch.no_grad():
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, self.warmup)):    # also allocates the carried state buffers
                    self._run_first()
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._first_out = self._run_first()
            graph.replay()
        self._first_graph = graph
        self._first_keys = self._mixer_keys()
        self._first_weights = [mod._kernel_weights() for mod in self._mixers()]

    @staticmethod
    def _as_list(state) -> List[LayerState]:
        if isinstance(state, dict):
            return [state[i] for i in range(len(state))]
        return list(state)

    def _mixers(self):
        return [mod for mod in self.model.modules() if hasattr(mod, "weights_key")]

    def _mixer_keys(self) -> list:
        return [mod.weights_key() for mod in self._mixers()]

    def _temporal_rows(self, t_tokens: int, x: Tensor) -> Tensor:
        m = self.model
        m._temporal_rows_override = None
        dtype = next(m.parameters()).dtype
        return m._get_temporal_pos_embedding(t_tokens, offset=self.offset, dtype=dtype,
                                             device=x.device).contiguous()

    def _run(self):
        m = self.model
        out = m(self._x, ssm_state=self._state, temporal_pos_offset=max(self.offset, 1))
        for (c, s), (nc, ns) in zip(self._state, self._as_list(out[-1])):
            c.copy_(nc)
            s.copy_(ns)
        return out

    def _capture(self, x: Tensor, rows: Tensor) -> None:
        m = self.model
        self._x = x.clone()
        self._rows = rows.clone()
        keep = [(c.clone(), s.clone()) for c, s in self._state]
        m._temporal_rows_override = self._rows
        try:
            with torch.no_grad():
                side = torch.cuda.Stream(device=x.device)
                side.wait_stream(torch.cuda.current_stream(x.device))
                with torch.cuda.stream(side):
                    for _ in range(self.warmup):        # warm the allocator and the tensor-map caches
                        self._run()
                torch.cuda.current_stream(x.device).wait_stream(side)
                for (c, s), (kc, ks) in zip(self._state, keep):   # warm-up advanced the state: restore
                    c.copy_(kc)
                    s.copy_(ks)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._out = self._run()
                for (c, s), (kc, ks) in zip(self._state, keep):
                    c.copy_(kc)
                    s.copy_(ks)
                graph.replay()
            self._graph = graph
            self._captured_keys = self._mixer_keys()
            self._captured_weights = [mod._kernel_weights() for mod in self._mixers()]
        finally:
            m._temporal_rows_override = None
