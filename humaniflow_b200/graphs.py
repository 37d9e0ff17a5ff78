"""CUDA-graph execution of fixed-shape steps of the hot path.

One pass of the path is ~75 kernel launches (49 convolutions, pooling, heads, the flow chain, LBS); issued one by one from
Python they keep a host thread busy for most of the ~1.5 ms the GPU needs, and the programmatic-dependent-launch edges between
the convolutions only help if the next launch is already queued.  ``CudaGraphRunner`` records the launches of a callable once
(all kernels of this library launch on torch's current stream, so they are capturable; the PDL edges are kept as programmatic
graph edges) and replays them with one driver call.

    step = CudaGraphRunner(lambda x, z, se: predict(model, smpl, x, z, se), x0, z0, se0)
    out = step(x, z, se)          # inputs are copied into the captured buffers, outputs are the captured tensors

The reference has nothing comparable (eager PyTorch); this is part of the new design ("CUDA streams and graphs").
Static shapes only: build one runner per (B, N, H, W).  Outputs are overwritten by the next replay.
"""
import torch


class CudaGraphRunner:
    def __init__(self, fn, *example_inputs, warmup=3):
        """fn(*tensors) -> tensor / tuple / dict of tensors.  ``example_inputs`` fix shapes, dtypes and device; non-tensor
        arguments are passed through unchanged on every replay."""
        self.fn = fn
        self.static_in = [a.clone() if torch.is_tensor(a) else a for a in example_inputs]
        dev = next(a.device for a in self.static_in if torch.is_tensor(a))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):               # packs weights, sizes workspaces, builds launch plans: none of that may happen under capture
                fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for s, a in zip(self.static_in, inputs):
            if torch.is_tensor(s) and a is not None and a is not s:
                s.copy_(a, non_blocking=True)
        self.graph.replay()
        return self.static_out


def predict_step(model, smpl, x, base_noise, shape_eps):
    """image (or StagedInput) -> N sampled meshes per image: what predict_humaniflow.py:112-160 does per batch.
    Returns (model output dict, vertices (B*N,V,3), joints (B*N,90,3))."""
    B, N = base_noise.shape[:2]
    out = model(x, num_samples=N, base_noise=base_noise, shape_eps=shape_eps)
    R = out['pose_rotmats_samples'].view(B * N, -1, 3, 3)
    so = smpl.forward_samples(out['shape_samples'].view(B * N, -1), R, out['glob_rotmat'], N)
    return out, so.vertices, so.joints
