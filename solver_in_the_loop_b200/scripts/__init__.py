"""CLI drop-ins for the reference's scripts (same flags)."""

def device_index(gpu_flag: str) -> int:
    """`--gpu` of the reference scripts is a CUDA_VISIBLE_DEVICES string; its Makefiles pass "-1" (= CPU) for data generation and
    apply.  This engine has no CPU path: "-1" / "" select device 0, "2,3" selects 2."""
    first = str(gpu_flag).split(",")[0].strip()
    try:
        idx = int(first)
    except ValueError:
        idx = 0
    if idx < 0:
        import logging
        logging.getLogger("solver_in_the_loop_b200").warning(
            '--gpu "%s" asks the reference for its CPU path; this engine has none: running on CUDA device 0', gpu_flag)
    return idx if idx >= 0 else 0


def reject_unimplemented(p: dict, names):
    """Flags the reference implements and this package does not must not vanish silently."""
    for n in names:
        if p.get(n):
            raise SystemExit("--%s is not implemented by solver_in_the_loop_b200 (the reference's %s path needs its Keras/TF "
                             "checkpoint machinery); remove the flag" % (n.replace("_", "-"), n))


def rank0_preprocess_then_barrier(make_dataset, skip_ds: bool, rank: int, world: int):
    """Down-sampling writes ds_*.npz next to the data: only rank 0 does it, the others wait and then only read."""
    import torch.distributed as dist
    if world > 1 and rank != 0:
        dist.barrier()
        return make_dataset(True)
    ds = make_dataset(skip_ds)
    if world > 1:
        dist.barrier()
    return ds


def load_init_weights(trainer, path, cin0: int, model: str):
    """--inittf: warm start from Keras-ordered weights (model.npz; model.h5 is mapped to the .npz beside it)."""
    from ..phi_compat import CorrectionModel
    m = CorrectionModel.load(path, cin0=cin0, device=trainer.weights.device, model=model)
    if m.flat.numel() != trainer.weights.numel():
        raise SystemExit("--inittf %s holds %d parameters, the model has %d" % (path, m.flat.numel(), trainer.weights.numel()))
    trainer.weights.copy_(m.flat)
