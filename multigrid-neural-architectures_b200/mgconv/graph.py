"""CUDA-graph replay of a whole training step.

The CIFAR / MNIST multigrid networks spend their time at 4x4 ... 1x1 grids: ~750 kernels of a few
microseconds each, so the step is bound by the host's launch rate (hard part 4).
Every libmgconv call is enqueued on the stream bound to the context and none synchronises, so a full
`zeroGradParameters -> ftrain -> btrain` step can be stream-captured once and replayed as one graph
launch.  torch is used for the capture plumbing only (torch.cuda.CUDAGraph = cudaStreamBeginCapture /
cudaGraphInstantiate / cudaGraphLaunch).

Constraints of a captured step: fixed shapes and buffers (inputs are copied INTO the captured tensors),
constant hyper-parameters (re-capture after NET.trainRule changes LR / WD), no host reads inside the step
(read the loss tensor after replay), single process (the NCCL bucket all-reduce of multigpu.py is left
un-captured).
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, warmup=2):
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):        # sizes every lazily grown workspace before capture
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            step_fn()

    def __call__(self):
        self.graph.replay()
