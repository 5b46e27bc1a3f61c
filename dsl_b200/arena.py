"""Chunked device arena for the plans' static buffers.

A plan (FCOSNet / TeacherPost / DSLEngine) allocates several hundred zero-initialised buffers once per input shape. As
separate `torch.zeros` calls that is one fill kernel each (~1000 launches per engine: slow for the small maps of the
multi-scale path, and it buries the step's own kernels in any launch trace). Here the buffers are carved out of a few
large chunks, each cleared by ONE memset when it is created; chunk sizes double from 64 MiB to 1 GiB, so a small plan is
one fill and the 9 GB benchmark plan about a dozen. Views keep their chunk alive; dropping the plan frees everything.
"""
import math

import torch

_ALIGN = 256   # bytes; TMA global addresses need 16, the swizzled tiles like 128


class Arena:
    def __init__(self, device, first_chunk=64 << 20, max_chunk=1 << 30):
        self.device = torch.device(device)
        self.next_chunk, self.max_chunk = int(first_chunk), int(max_chunk)
        self.chunk, self.off = None, 0
        self.chunks = 0
        self.bytes = 0

    def zeros(self, *shape, dtype=torch.float32):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        n = int(math.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if self.device.type != "cuda" or n == 0:
            return torch.zeros(*shape, dtype=dtype, device=self.device)
        need = (n + _ALIGN - 1) // _ALIGN * _ALIGN
        if self.chunk is None or self.off + need > self.chunk.numel():
            size = max(self.next_chunk, need)
            self.next_chunk = min(self.next_chunk * 2, self.max_chunk)
            self.chunk = torch.zeros(size, dtype=torch.uint8, device=self.device)
            self.off = 0
            self.chunks += 1
            self.bytes += size
            # cudaMalloc returns >= 256-byte aligned blocks, torch's caching allocator 512
            assert self.chunk.data_ptr() % _ALIGN == 0
        v = self.chunk[self.off:self.off + n].view(dtype).view(*shape)
        self.off += need
        return v
