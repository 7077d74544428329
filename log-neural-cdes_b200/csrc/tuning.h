// Kernel-variant switches (internal).  Every variant computes the same result; which one is faster is a
// hardware question answered by measurement, so the choice is a process-wide tuning value settable through
// lncde_set_tuning() (include/lncde_b200.h) or the environment (LNCDE_TUNE_<KEY>=0/1 read at first use).
#pragma once

namespace lncde {

enum TuneKey {
  kTuneK1Tma = 0,    // "k1_tma":   closed-form logsig kernel with TMA bulk-copy staging (logsig_levy_tma_kernel)
  kTuneK3Fused = 1,  // "k3_fused": specialised Log-NCDE kernels with the 32 x 32 layer phases fused on single warps
  kTuneK2FusedD = 2, // "k2_fused_d": read by the host mirror - diagonal LNCDE through the fused K2d kernels (dlncde.cu)
  kTuneK1GridCap = 3, // "k1_grid_cap": upper bound on the grid of the persistent K1 kernel (0 = machine-sized); experiments, tests
  kTuneK2DeV2 = 4,   // "k2_de_v2": big-block (DE) LNCDE - 8 x 8 register-tile flow build, reverse scan stores lam_t only and the
                     //             bracket gradient is a triple product of the small operands (linear_de.cuh)
  kTuneK3Stream = 5, // "k3_stream": generic Log-NCDE kernels - layers that do not fit in shared memory are read with the streamed
                     //              (prefetching, all-lanes) mat-vecs of logncde_common.cuh
  kTuneK2Pscan = 6,  // "k2_pscan": small-block (D / BD) LNCDE scans parallel in time - chunk products, boundary recurrence over the
                     //             chunks, per-chunk sweeps (linear_pscan.cuh) instead of T - 1 dependent steps per sample
  kTuneCount
};

int tuning_value(int key);

}  // namespace lncde
