"""The CLI drop-ins accept the reference Makefiles' command lines verbatim (flag names AND meanings): every command line of
karman-2d/Makefile and burgers/Makefile that invokes an in-scope script is fed through the corresponding parse()."""
import shlex

import pytest

from solver_in_the_loop_b200.scripts import burgers, burgers_apply, burgers_train, karman, karman_apply, karman_train

# (module, command line after the script name, {dest: expected value}) — shell arithmetic / printf already expanded for i = 0
KARMAN = [
    # karman-2d/Makefile:22 (karman-fdt-hires-set)
    (karman, '-o karman-fdt-hires-set -r 128 -l 100 --re 160000 --gpu "-1" --seed 0 --thumb',
     dict(res=128, len=100, re=160000.0, skipsteps=999, simsteps=1500, scale=4)),
    # karman-2d/Makefile:36-38 (karman-fdt-lores-set): --skipsteps 0 -t 500 -d 4 --initdH/--initvH
    (karman, '-o karman-fdt-lores-set -r 32  -l 100 --re 160000 --gpu "-1" --seed 0 --thumb --skipsteps 0 -t 500 -d 4 '
             '--initdH karman-fdt-hires-set/sim_000000/dens_001000.npz --initvH karman-fdt-hires-set/sim_000000/velo_001000.npz',
     dict(res=32, skipsteps=0, simsteps=500, scale=4, initdH="karman-fdt-hires-set/sim_000000/dens_001000.npz",
          initvH="karman-fdt-hires-set/sim_000000/velo_001000.npz")),
    # karman-2d/Makefile:74-75 (NON) and :79-80 (SOL-32)
    (karman_train, "--tf karman-fdt-non/tf --log karman-fdt-non/tf/run.log --epochs=100 --lr 0.0001 -l 100 -t 500 -s 4 -m 1 -n 6 -b 3 --seed 0 "
                   "--gpu '0' --cuda --train karman-fdt-hires-set", dict(msteps=1, nsims=6, sbatch=3, scale=4, simsteps=500, epochs=100, lr=1e-4)),
    (karman_train, "--tf karman-fdt-sol32/tf --log karman-fdt-sol32/tf/run.log --epochs=100 --lr 0.0001 -l 100 -t 500 -s 4 -m 32 -n 6 -b 3 --seed 0 "
                   "--gpu '0' --cuda --train karman-fdt-hires-set", dict(msteps=32, nsims=6, sbatch=3, train="karman-fdt-hires-set")),
    # karman-2d/Makefile:121-127 (karman-fdt-sol32/run_test)
    (karman_apply, '-o karman-fdt-sol32/run_test --stats karman-fdt-sol32/tf/dataStats.pickle --model karman-fdt-sol32/tf/model.h5 --gpu "-1" '
                   "--initdH karman-fdt-hires-testset/sim_000000/dens_001000.npz --initvH karman-fdt-hires-testset/sim_000000/velo_001000.npz "
                   "-s 4 -r 32 -l 100 --re 240000 -t 500", dict(scale=4, res=32, re=240000.0, simsteps=500, model="karman-fdt-sol32/tf/model.h5")),
]
BURGERS = [
    # burgers/Makefile:22 (hires set), :37-40 (lores set)
    (burgers, '-o burgers-fdt-hires-set -r 128 -l 32 --dt 0.1 --skipsteps 30 -t 200 --gpu "-1" --seed 0 --thumb', dict(res=128, len=32, dt=0.1, skipsteps=30, simsteps=200)),
    (burgers, '-o burgers-fdt-lores-set -r 32 -l 32 --dt 0.1 --skipsteps 0 -t 200 -d 4 --seed 0 --initvH burgers-fdt-hires-set/sim_000000/velo_000000.npz '
              '--loadfH "burgers-fdt-hires-set/sim_000000/forc_0*.npz" --gpu "-1" --thumb', dict(scale=4, skipsteps=0, loadfH="burgers-fdt-hires-set/sim_000000/forc_0*.npz")),
    # burgers/Makefile:75-77 (SOL-04)
    (burgers_train, "--tf burgers-fdt-sol04/tf --log burgers-fdt-sol04/tf/run.log --epochs 100 --lr 0.0001 -l 32 --dt 0.1 -t 200 -s 4 -m 4 -n 10 -b 5 --seed 0 "
                    "--gpu '0' --cuda --train burgers-fdt-hires-set", dict(msteps=4, nsims=10, sbatch=5, dt=0.1, len=32)),
    # burgers/Makefile:104-110 (burgers-fdt-sol04/run_test)
    (burgers_apply, '-o burgers-fdt-sol04/run_test --stats burgers-fdt-sol04/tf/dataStats.pickle --model burgers-fdt-sol04/tf/model.h5 --gpu "-1" '
                    '--initvH burgers-fdt-hires-testset/sim_000000/velo_000000.npz --loadfH "burgers-fdt-hires-testset/sim_000000/forc_0*.npz" '
                    "-s 4 -r 32 -l 32 --dt 0.1 -t 200", dict(scale=4, res=32, len=32, dt=0.1, simsteps=200)),
]


@pytest.mark.parametrize("mod,line,expect", KARMAN + BURGERS, ids=lambda v: getattr(v, "__name__", None))
def test_reference_makefile_command_lines_parse(mod, line, expect):
    p = vars(mod.parse(shlex.split(line)))
    for k, v in expect.items():
        assert p[k] == v, (k, p[k], v)


def test_reference_defaults():
    """Defaults the reference scripts rely on when a flag is omitted (karman.py:34-46, karman_apply.py:22-32, burgers_train.py:22-44)."""
    p = vars(karman.parse([]))
    assert (p["skipsteps"], p["scale"], p["res"], p["simsteps"]) == (999, 4, 32, 1500)
    p = vars(karman_apply.parse([]))
    assert p["output"] == "/tmp/phiflow/run" and p["stats"] == "/tmp/phiflow/data/dataStats.pickle" and p["model"] == "/tmp/phiflow/tf/model.h5"
    p = vars(burgers_train.parse([]))
    assert (p["len"], p["nsims"], p["sbatch"], p["seed"], p["simsteps"], p["msteps"]) == (32, 10, 2, 0, 200, 2)
    p = vars(karman_train.parse([]))
    assert (p["scale"], p["nsims"], p["sbatch"], p["simsteps"], p["msteps"], p["len"], p["lr"]) == (4, 1, 1, 1500, 2, 100, 1e-3)


@pytest.mark.parametrize("mod,flag", [(karman_train, "--pretf"), (karman_train, "--reg-loss"), (burgers_train, "--pretf")])
def test_unimplemented_flags_raise(mod, flag):
    """Flags the reference implements and this package does not (supervised-baseline weights, L2 regulariser) must not vanish silently."""
    from solver_in_the_loop_b200.scripts import reject_unimplemented
    argv = [flag] + (["x.h5"] if flag == "--pretf" else [])
    p = vars(mod.parse(argv))
    with pytest.raises(SystemExit):
        reject_unimplemented(p, ("pretf", "reg_loss"))
