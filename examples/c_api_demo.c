/* The C ABI without any Python: 2-D Mueller-Brown Langevin (BASELINE C2 physics) for M walkers.
 *
 *   gcc -O2 -I include examples/c_api_demo.c -o c_api_demo \
 *       -L ndsimulator_b200/csrc -lndsb200 -Wl,-rpath,$PWD/ndsimulator_b200/csrc -lm
 *   ./c_api_demo [walkers] [steps]
 *
 * Prints the mean temperature, which the reference's first-order Langevin integrator (same xi in both half
 * kicks) puts at T (1 + c1)^2 / (1 + c1^2) = 599.6 K for T = 300 K, gamma dt = 0.1.            */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ndsb200.h"

int main(int argc, char** argv) {
  const int64_t M = argc > 1 ? atoll(argv[1]) : 65536;
  const int64_t steps = argc > 2 ? atoll(argv[2]) : 5000;
  nds_config cfg;
  nds_config_defaults(&cfg);
  cfg.precision = NDS_F32;
  cfg.ndim = 2;
  cfg.potential = NDS_POT_MUELLER2D;
  cfg.integrator = NDS_INT_LANGEVIN;
  cfg.gamma = 0.1; cfg.dt = 1.0; cfg.temperature = 300.0; cfg.initial_temperature = 300.0; cfg.mass = 1.0;
  cfg.n_walkers = M; cfg.n_walkers_global = M; cfg.walker_offset = 0;
  cfg.steps_per_launch = 1000;
  cfg.seed = 2024;
  nds_engine* e = NULL;
  if (nds_create(&cfg, &e)) { fprintf(stderr, "nds_create: %s\n", nds_last_error(NULL)); return 1; }
  double* x = (double*)malloc(sizeof(double) * 2 * M);
  double* thermo = (double*)malloc(sizeof(double) * 5 * M);
  for (int64_t w = 0; w < M; w++) { x[w] = 22.0; x[M + w] = 24.0; }
  if (nds_set_state(e, x, NULL) || nds_begin(e) || nds_run(e, steps) || nds_sync(e) ||
      nds_get_state(e, x, NULL, NULL, NULL, NULL, thermo)) {
    fprintf(stderr, "engine: %s\n", nds_last_error(e));
    return 1;
  }
  double T = 0.0, pe = 0.0;
  for (int64_t w = 0; w < M; w++) { T += thermo[w]; pe += thermo[M + w]; }
  printf("kernel %s\nwalkers %lld steps %lld launches %lld last launch block %.3f ms\n<T> = %.2f K  <pe> = %.5f eV\n",
         nds_kernel_name(e), (long long)M, (long long)nds_step(e), (long long)nds_launch_count(e),
         nds_last_run_ms(e), T / M, pe / M);
  nds_destroy(e);
  free(x); free(thermo);
  return 0;
}
