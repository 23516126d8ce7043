"""CUDA-graph replay of a whole forward pass (CUDA streams and graphs instead of a tracing compiler).

The forward passes here are static launch sequences (DenseNet: 104 kernels, SPADE generator: ~435) whose host-side cost
(Python + ctypes, ~10-15 us per launch) exceeds the kernels' run time at small batch.  `graphed_call` captures the
sequence once per (input shapes, parameter versions) into a torch.cuda.CUDAGraph -- the C-ABI launches go to torch's current
(capturing) stream, workspaces come from the graph's private pool -- and replays it afterwards with the inputs copied into
the captured buffers.  Outputs are returned as fresh tensors (one device copy) so callers may keep them.
"""
import torch


def graphed_call(cache, state_key, fn, inputs):
    key = (state_key,) + tuple((tuple(t.shape), t.dtype, str(t.device)) for t in inputs)
    entry = cache.get(key)
    if entry is None:
        cache.clear()                                   # parameters or shapes changed: drop stale graphs (and their memory)
        static_in = [t.clone() for t in inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                   # warm-up outside capture: packs weights, builds tables, sizes workspaces
            fn(*static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = fn(*static_in)
        entry = (graph, static_in, static_out)
        cache[key] = entry
    graph, static_in, static_out = entry
    for s, t in zip(static_in, inputs):
        s.copy_(t)
    graph.replay()
    if isinstance(static_out, (list, tuple)):
        return type(static_out)(o.clone() for o in static_out)
    return static_out.clone()
