#!/usr/bin/env python
"""Build oracle/_ref/libhamers_ref.so from the reference's OWN source files (test infrastructure).

The HAMeRS build needs MPI + HDF5 + SAMRAI + Fortran, none of which exist in this image, so
the reference cannot be built as a whole.  Its per-point `static inline` kernels, however, are
SAMRAI-free.  This recipe reads them where they lie under /root/reference, wraps them in
namespaces inside a generated translation unit under oracle/_ref/ (git-ignored, never
committed), compiles it with the reference's own flags (g++ -std=c++11 -O3, CMakeLists.txt:80;
HAMERS_EPSILON = 1.0e-15, include/HAMeRS_config.hpp.in:16) and deletes the generated source.
The resulting library pins the oracle's WCNS5-JS / WCNS5-Z / WCNS6-LD interpolations and HLLC / HLLC-HLL kernels
(tests/test_oracle_pinned.py).

Nothing here is copied into the repository; if /root/reference is absent (the GPU box) the
script does nothing and the prebuilt .so that travelled with the snapshot is used.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HAMERS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

SOURCES = {
    # namespace -> reference file holding `static inline` point kernels
    "ref_weno": "src/flow/convective_flux_reconstructors/WCNS56/ConvectiveFluxReconstructorWCNS5-JS-HLLC-HLL.cpp",
    "ref_weno_z": "src/flow/convective_flux_reconstructors/WCNS56/ConvectiveFluxReconstructorWCNS5-Z-HLLC-HLL.cpp",
    "ref_weno_ld": "src/flow/convective_flux_reconstructors/WCNS56/ConvectiveFluxReconstructorWCNS6-LD-HLLC-HLL.cpp",
    "ref_ss_hllc": "src/flow/flow_models/single-species/Riemann_solvers/FlowModelRiemannSolverSingleSpeciesHLLC.cpp",
    "ref_ss_hyb": "src/flow/flow_models/single-species/Riemann_solvers/FlowModelRiemannSolverSingleSpeciesHLLC-HLL.cpp",
    "ref_fe_hllc": "src/flow/flow_models/five-eqn_Allaire/Riemann_solvers/FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp",
    "ref_fe_hyb": "src/flow/flow_models/five-eqn_Allaire/Riemann_solvers/FlowModelRiemannSolverFiveEqnAllaireHLLC-HLL.cpp",
}


EOS_SOURCE = "src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateIdealGas.cpp"
EOS_SCALARS = ("getPressure", "getSoundSpeed", "getInternalEnergy")


def member_function(text: str, cls: str, name: str) -> str:
    """The FIRST definition `double\n<cls>::<name>(...) const { ... }` in text (the scalar overload comes first in the
    reference's file), as text."""
    m = re.search(r"^double\s*\n" + cls + "::" + name + r"\(", text, flags=re.M)
    start = m.start()
    brace = text.index("{", m.end())
    depth, i = 0, brace
    while True:
        ch = text[i]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return text[start:i + 1]


def statement(text: str, pattern: str, occurrence: int = -1) -> str:
    """The C statement of `text` that starts at the given occurrence of regex `pattern` and runs to the next ';'."""
    ms = list(re.finditer(pattern, text))
    m = ms[occurrence]
    return text[m.start():text.index(";", m.end()) + 1]


def path_statements() -> str:
    """A function built around the reference's OWN statements for the first derivative (DerivativeFirstOrder.cpp:601),
    dilatation and vorticity magnitude (ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:1631, 1653-1657), the sensor
    value (:2098-2101) and the face flux (:2370-2375): the statements are read from /root/reference and compiled
    verbatim; only the scaffolding around them (arrays of length 1-3, index constants) is ours."""
    with open(os.path.join(REF, "src/util/derivatives/DerivativeFirstOrder.cpp")) as fh:
        der = fh.read()
    with open(os.path.join(REF, SOURCES["ref_weno"].replace("WCNS5-JS-HLLC-HLL", "WCNS56-HLLC-HLL"))) as fh:
        w56 = fh.read()
    s_der = statement(der, r"dudx\[idx_derivative\] = \(double\(1\)/double\(2\)")
    s_theta = statement(w56, r"theta\[idx\] = dudx\[idx\] \+ dvdy\[idx\] \+ dwdz\[idx\]")
    s_ox = statement(w56, r"const double omega_x = dwdy\[idx\]")
    s_oy = statement(w56, r"const double omega_y = dudz\[idx\]")
    s_oz = statement(w56, r"const double omega_z = dvdx\[idx\] - dudy\[idx\]")
    s_Om = statement(w56, r"Omega\[idx\] = sqrt\(omega_x")
    s_ta = statement(w56, r"double theta_avg = 0\.5\*\(theta\[idx_L\]")
    s_Oa = statement(w56, r"double Omega_avg = 0\.5\*\(Omega\[idx_L\]")
    s_s = statement(w56, r"s_x\[idx_midpoint_x\] = -theta_avg")
    s_F = statement(w56, r"F_face_x\[idx_face_x\] = dt\*\(")
    return f"""
extern "C" void ref_path_points(const double in[16], double out[5])
{{
    {{
        const double u[2] = {{in[1], in[0]}};
        const int idx_x_L = 0, idx_x_R = 1, idx_derivative = 0;
        const double dx = in[2];
        double dudx[1];
        {s_der}
        out[0] = dudx[0];
    }}
    double theta[2], Omega[2];
    {{
        const int idx = 0;
        const double dudx[1] = {{in[3]}}, dudy[1] = {{in[4]}}, dudz[1] = {{in[5]}}, dvdx[1] = {{in[6]}}, dvdy[1] = {{in[7]}},
                     dvdz[1] = {{in[8]}}, dwdx[1] = {{in[9]}}, dwdy[1] = {{in[10]}}, dwdz[1] = {{in[11]}};
        {s_theta}
        {s_ox}
        {s_oy}
        {s_oz}
        {s_Om}
        out[1] = theta[0];
        out[2] = Omega[0];
    }}
    {{
        theta[1] = 0.75*out[1] - in[3];
        Omega[1] = 1.25*out[2];
        const int idx_L = 0, idx_R = 1, idx_midpoint_x = 0;
        double s_x[1];
        {s_ta}
        {s_Oa}
        {s_s}
        out[3] = s_x[0];
    }}
    {{
        const double dt = in[12];
        double mid[3] = {{in[13], in[14], in[15]}}, node[2] = {{in[1], in[0]}};
        double* F_midpoint_x[1] = {{mid}};
        double* F_node_x[1] = {{node}};
        const int ei = 0, idx_midpoint_x_L = 0, idx_midpoint_x = 1, idx_midpoint_x_R = 2, idx_node_L = 0, idx_node_R = 1, idx_face_x = 0;
        double F_face_x[1];
        {s_F}
        out[4] = F_face_x[0];
    }}
}}
"""


def static_inline_functions(text: str) -> str:
    """Return the concatenation of every `static inline ...` function definition in text."""
    out = []
    for m in re.finditer(r"^static inline[^\n]*\n?[^\n;{]*\(", text, flags=re.M):
        start = m.start()
        brace = text.index("{", m.end())
        depth, i = 0, brace
        while True:
            ch = text[i]
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        out.append(text[start:i + 1])
    return "\n\n".join(out)


WRAPPERS = r"""
extern "C" {

/* WCNS5-JS point interpolation: performLocalWENOInterpolationMinus/Plus with idx_side = 0 */
void ref_weno5js_point(const double U[6], int p, double* U_minus, double* U_plus)
{
    double vals[6]; double* Ua[6];
    for (int m = 0; m < 6; m++) { vals[m] = U[m]; Ua[m] = &vals[m]; }
    ref_weno::performLocalWENOInterpolationMinus(U_minus, Ua, 0, p);
    ref_weno::performLocalWENOInterpolationPlus(U_plus, Ua, 0, p);
}

/* WCNS5-Z and WCNS6-LD point interpolation (SURVEY row f2) */
void ref_weno5z_point(const double U[6], int p, double* U_minus, double* U_plus)
{
    double vals[6]; double* Ua[6];
    for (int m = 0; m < 6; m++) { vals[m] = U[m]; Ua[m] = &vals[m]; }
    ref_weno_z::performLocalWENOInterpolationMinus(U_minus, Ua, 0, p);
    ref_weno_z::performLocalWENOInterpolationPlus(U_plus, Ua, 0, p);
}

void ref_weno6ld_point(const double U[6], int p, int q, double C, double alpha_tau, double* U_minus, double* U_plus)
{
    double vals[6]; double* Ua[6];
    for (int m = 0; m < 6; m++) { vals[m] = U[m]; Ua[m] = &vals[m]; }
    ref_weno_ld::performLocalWENOInterpolationMinus(U_minus, Ua, 0, p, q, C, alpha_tau);
    ref_weno_ld::performLocalWENOInterpolationPlus(U_plus, Ua, 0, p, q, C, alpha_tau);
}

/* HLLC and HLLC-HLL point kernels on one face (idx = idx_flux = 0).  The midpoint normal velocity is
 * the two-line driver formula (SingleSpeciesHLLC.cpp:3381-3388, FiveEqnAllaireHLLC.cpp:6034-6038). */
int ref_riemann_point(int model, int dim, int ns, int dir,
                      const double* V_L, const double* V_R,
                      double rho_L, double rho_R, double c_L, double c_R, double eps_L, double eps_R,
                      double* F_HLLC, double* F_HYB, double* vel_mid)
{
    const int neq = model == 0 ? dim + 2 : dim + 2*ns;
    double vl[16], vr[16], f1[16], f2[16];
    double *VL[16], *VR[16], *F1[16], *F2[16];
    for (int e = 0; e < neq; e++) { vl[e] = V_L[e]; vr[e] = V_R[e]; VL[e] = &vl[e]; VR[e] = &vr[e]; F1[e] = &f1[e]; F2[e] = &f2[e]; }
    double s_minus = 0, s_plus = 0, s_star = 0, Chi = 0;
    double s_minus2 = 0, s_plus2 = 0, s_star2 = 0, Chi2 = 0;
    const int key = model*100 + dim*10 + dir;
    switch (key) {
    case  20: ref_ss_hllc::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC2D(F1, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0);
              ref_ss_hyb::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC_HLL2D(F2, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0); break;
    case  21: ref_ss_hllc::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC2D(F1, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0);
              ref_ss_hyb::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC_HLL2D(F2, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0); break;
    case  30: ref_ss_hllc::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC3D(F1, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0);
              ref_ss_hyb::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC_HLL3D(F2, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0); break;
    case  31: ref_ss_hllc::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC3D(F1, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0);
              ref_ss_hyb::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC_HLL3D(F2, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0); break;
    case  32: ref_ss_hllc::computeLocalConvectiveFluxInZDirectionFromPrimitiveVariablesHLLC3D(F1, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0);
              ref_ss_hyb::computeLocalConvectiveFluxInZDirectionFromPrimitiveVariablesHLLC_HLL3D(F2, VL, VR, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0); break;
    case 120: ref_fe_hllc::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC2D(F1, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0, ns, neq);
              ref_fe_hyb::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC_HLL2D(F2, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0, ns, neq); break;
    case 121: ref_fe_hllc::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC2D(F1, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0, ns, neq);
              ref_fe_hyb::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC_HLL2D(F2, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0, ns, neq); break;
    case 130: ref_fe_hllc::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC3D(F1, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0, ns, neq);
              ref_fe_hyb::computeLocalConvectiveFluxInXDirectionFromPrimitiveVariablesHLLC_HLL3D(F2, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0, ns, neq); break;
    case 131: ref_fe_hllc::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC3D(F1, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0, ns, neq);
              ref_fe_hyb::computeLocalConvectiveFluxInYDirectionFromPrimitiveVariablesHLLC_HLL3D(F2, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0, ns, neq); break;
    case 132: ref_fe_hllc::computeLocalConvectiveFluxInZDirectionFromPrimitiveVariablesHLLC3D(F1, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0, ns, neq);
              ref_fe_hyb::computeLocalConvectiveFluxInZDirectionFromPrimitiveVariablesHLLC_HLL3D(F2, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus2, s_plus2, s_star2, Chi2, 0, 0, ns, neq); break;
    default: return -1;
    }
    for (int e = 0; e < neq; e++) { F_HLLC[e] = f1[e]; F_HYB[e] = f2[e]; }
    const int iun = (model == 0 ? 1 : ns) + dir;
    if (s_star > double(0)) *vel_mid = vl[iun] + s_minus*(Chi - double(1));
    else                    *vel_mid = vr[iun] + s_plus*(Chi - double(1));
    return 0;
}

}
"""


EOS_WRAPPER = r"""
extern "C" void ref_eos_point(double gamma, double rho, double epsilon, double* p, double* c, double* eps_back)
{
    ref_eos::EquationOfStateIdealGas eos;
    std::vector<const double*> thermo(1, &gamma);
    *p = eos.getPressure(&rho, &epsilon, thermo);
    *c = eos.getSoundSpeed(&rho, p, thermo);
    *eps_back = eos.getInternalEnergy(&rho, p, thermo);
}
"""


def main() -> int:
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: keeping any prebuilt oracle/_ref/libhamers_ref.so")
        return 0
    os.makedirs(OUT, exist_ok=True)
    parts = ["#include <cmath>\n#include <algorithm>\n#define HAMERS_EPSILON 1.0e-15\n#define EPSILON HAMERS_EPSILON\n"]
    for ns, rel in SOURCES.items():
        with open(os.path.join(REF, rel)) as fh:
            body = static_inline_functions(fh.read())
        parts.append(f"namespace {ns} {{\n{body}\n}}\n")
    # the scalar ideal-gas EOS members (EquationOfStateIdealGas.cpp:29-45, 561-577, 1093-1108) behind a stub class
    with open(os.path.join(REF, EOS_SOURCE)) as fh:
        eos_text = fh.read()
    decls = "\n".join(f"    double {n}(const double* const, const double* const, const std::vector<const double*>&) const;"
                      for n in EOS_SCALARS)
    bodies = "\n\n".join(member_function(eos_text, "EquationOfStateIdealGas", n) for n in EOS_SCALARS)
    parts.append("#include <vector>\nnamespace ref_eos {\nstruct EquationOfStateIdealGas {\n" + decls + "\n};\n" + bodies + "\n}\n")
    parts.append(WRAPPERS)
    parts.append(EOS_WRAPPER)
    parts.append(path_statements())
    gen = os.path.join(OUT, "_generated_ref_kernels.cpp")
    with open(gen, "w") as fh:
        fh.write("\n".join(parts))
    so = os.path.join(OUT, "libhamers_ref.so")
    cmd = ["g++", "-std=c++11", "-O3", "-Wno-deprecated", "-Wno-unused-function", "-fPIC", "-shared", "-o", so, gen]
    try:
        subprocess.check_call(cmd)
    finally:
        os.remove(gen)   # reference text never stays on disk outside /root/reference
    print(f"[build_ref] built {so}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
