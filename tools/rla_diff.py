"""Buffer-by-buffer difference between the CUDA RLA_ResNet plan and its CPU emulation (tests/emu_lib.py): prints every
forward / backward buffer of every block with its relative L2 difference, so the first kernel that diverges is named.
Usage (GPU box): python tools/rla_diff.py [B H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from tests.test_gpu_rla import compare_backbones, _l2
    a = [int(v) for v in sys.argv[1:4]] or [2, 128, 192]
    rows, net, emu, *_ = compare_backbones(*a, verbose=False)
    for tag, k, e in rows:
        flag = "  <<<<" if (e > 3e-2 or (k.endswith((".wp", ".wpT")) and e > 1e-4)) else ""
        print(f"{tag:14s} {k:10s} {e:.3e}{flag}")
    print("flat gradient rel-L2", _l2(net.grad, emu.grad))
    for p in net.store.spec:
        if p.region != "F":
            e = _l2(net.grad_view(p.name), emu.grad_view(p.name))
            if e > 3e-2:
                print(f"grad {p.name:40s} {e:.3e}  <<<<")


if __name__ == "__main__":
    main()
