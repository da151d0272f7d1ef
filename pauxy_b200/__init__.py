"""pauxy_b200: B200-native phaseless-AFQMC walker propagation and local energy
for generic Cholesky Hamiltonians with a single-determinant trial -- the hot
path of pauxy-qmc/pauxy, batched over walkers in hand-written sm_100a CUDA
behind the reference's driver / propagator / walker / estimator interface."""

__version__ = '0.1.0'
