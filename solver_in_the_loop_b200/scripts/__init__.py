"""CLI drop-ins for the reference's scripts (same flags)."""

def device_index(gpu_flag: str) -> int:
    """`--gpu` of the reference scripts is a CUDA_VISIBLE_DEVICES string; its Makefiles pass "-1" (= CPU) for data generation and
    apply.  This engine has no CPU path: "-1" / "" select device 0, "2,3" selects 2."""
    first = str(gpu_flag).split(",")[0].strip()
    try:
        idx = int(first)
    except ValueError:
        idx = 0
    return idx if idx >= 0 else 0
