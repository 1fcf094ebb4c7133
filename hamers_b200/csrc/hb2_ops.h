/*
 * hb2_ops.h -- internal launch table shared by hb2_sweeps.cu (compiled twice: exact / fast
 * arithmetic) and hb2_abi.cu.
 */
#pragma once
#include <cuda_runtime.h>
#include "hb2_core.cuh"
#include "hb2_sensor.cuh"

namespace hb2 {

struct AdvanceArgs {
    Geom G;
    int model, ns, neq, ncomp;
    int ncoef;
    double alpha[HB2_MAXS], beta[HB2_MAXS], gamma[HB2_MAXS];
    const double* Uint[HB2_MAXS][HB2_MAXC];
    const double* Fint[HB2_MAXS][3 * HB2_MAXE];
    const double* Sint[HB2_MAXS][HB2_MAXE];
    double* Uout[HB2_MAXC];
    double* Facc[3 * HB2_MAXE]; /* may be null */
    double* Sacc[HB2_MAXE];
};

struct LaunchCfg {
    int model, dim, ns;
};

struct Ops {
    /* the per-face s > 0.65 decisions (one byte per cell) on cells -1..N+1 */
    int (*sensor)(const LaunchCfg&, const SensorArgs&, cudaStream_t);
    /* one direction sweep */
    int (*sweep)(const LaunchCfg&, int dir, const DirArgs&, cudaStream_t);
    /* Euler::advanceSingleStepOnPatch from materialised fluxes */
    int (*advance)(const AdvanceArgs&, cudaStream_t);
};

const Ops* ops_exact();    /* WCNS5-JS, reference operation order */
const Ops* ops_exact_z();  /* WCNS5-Z */
const Ops* ops_exact_ld(); /* WCNS6-LD */
const Ops* ops_fast();     /* WCNS5-JS, re-associated */
const Ops* ops_fast_z();   /* WCNS5-Z, re-associated */
const Ops* ops_fast_ld();  /* WCNS6-LD, re-associated (constant_p = 2, constant_q = 4) */

}  // namespace hb2
