"""Real-time deployment path: what `ImageAgent.run_step` (PMoE/autoagents/image_agent.py:127-177)
does every simulator tick — transform the new camera frame (crop / resize / ToTensor), push it into the window of the last
`n_frames` frames, build the (1, T, 3, H, W) / speed / one-hot command tensors and call the policy — as ONE captured CUDA
graph replay per tick: the upload of the raw uint8 frame from pinned memory, the Pillow-exact preprocessing kernels
(preproc.py), the window shift, every kernel of `model.sample(...)` and the download of the action.  The reference rebuilds
the tensors on the host each tick and launches several hundred small kernels eagerly; at B = 1 those launches, not the
arithmetic, are the latency.
"""
import numpy as np
import torch

from .preproc import FramePreprocessor


class RealtimeSampler:
    """policy(images (1,T,3,H,W), speed (1,1), command (1,n_commands)) -> action tensor; defaults to `model.sample`.

    step(rgb, speed, command_index) -> action on the host (torch.Tensor), `rgb` an (Hs, Ws, 3) uint8 RGB array.
    reset() empties the window: the next frame fills all T slots (a freshly spawned agent has seen one frame)."""

    def __init__(self, model, frame_hw, crop=(125, 90), resize=(224, 224), n_frames=4, n_commands=6, policy=None, graph=True,
                 device="cuda"):
        self.model = model.eval()
        self.policy = policy if policy is not None else model.sample
        self.dev = torch.device(device)
        self.T, self.n_commands = n_frames, n_commands
        hs, ws = frame_hw
        self.pp = FramePreprocessor(crop, resize, device)
        self.h_frame = torch.empty(1, hs, ws, 3, dtype=torch.uint8).pin_memory()
        self.h_scal = torch.zeros(1 + n_commands, dtype=torch.float32).pin_memory()     # [speed, one-hot command]
        self.d_frame = torch.empty(1, hs, ws, 3, dtype=torch.uint8, device=self.dev)
        self.d_scal = torch.zeros(1 + n_commands, dtype=torch.float32, device=self.dev)
        self.window = torch.zeros(1, n_frames, 3, resize[0], resize[1], dtype=torch.float32, device=self.dev)
        self.shifted = torch.zeros_like(self.window)
        self.h_action = None
        self.graph = None
        self.use_graph = graph
        self.fresh = True
        torch.distributions.Distribution.set_default_validate_args(False)  # argument validation synchronises the host

    def reset(self):
        self.fresh = True

    def _tick_body(self):
        self.d_frame.copy_(self.h_frame, non_blocking=True)
        self.d_scal.copy_(self.h_scal, non_blocking=True)
        # window <- window[1:] + new frame (two buffers: a device copy may not overlap itself)
        self.shifted[:, :-1].copy_(self.window[:, 1:])
        self.pp(self.d_frame, out=self.shifted[:, -1])
        self.window.copy_(self.shifted)
        with torch.no_grad():
            act = self.policy(self.window, self.d_scal[0:1].view(1, 1), self.d_scal[1:].view(1, -1))
        act = act.reshape(-1).float()
        if self.h_action is None:
            self.h_action = torch.empty(act.numel(), dtype=torch.float32).pin_memory()
        self.h_action.copy_(act, non_blocking=True)

    def _capture(self):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: lazy initialisation, weight-pack caches, allocator pool
                self._tick_body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._tick_body()
        self.graph = g

    def step(self, rgb, speed, command_index):
        frame = torch.from_numpy(np.ascontiguousarray(rgb)) if isinstance(rgb, np.ndarray) else rgb
        self.h_frame[0].copy_(frame)
        self.h_scal.zero_()
        self.h_scal[0] = float(speed)
        self.h_scal[1 + int(command_index)] = 1.0
        if self.fresh:  # first frame after a reset: every slot of the window shows it
            first = self.pp(self.h_frame.to(self.dev))
            self.window.copy_(first.view(1, 1, *first.shape[1:]).expand_as(self.window))
            self.fresh = False
        if self.use_graph:
            if self.graph is None:
                keep = self.window.clone()
                self._capture()           # the warm-up ticks shifted the window: restore it
                self.window.copy_(keep)
            self.graph.replay()
        else:
            self._tick_body()
        torch.cuda.current_stream().synchronize()
        return self.h_action.clone()
