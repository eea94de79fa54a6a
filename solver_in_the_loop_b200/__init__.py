"""solver_in_the_loop_b200 — B200-native engine for the unrolled differentiable solver step of
tum-pbs/Solver-in-the-Loop (karman-2d / burgers).  The compute path is libsol_b200.so (hand-written
sm_100a CUDA behind the C ABI in include/sol_b200.h); this package is the thin Python host layer
that mirrors the reference's simulator.step() / to_feature / to_staggered surface."""
__version__ = "0.1.0"
