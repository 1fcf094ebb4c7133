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
    # SURVEY row f3
    "ref_fc_hllc": "src/flow/flow_models/four-eqn_conservative/Riemann_solvers/FlowModelRiemannSolverFourEqnConservativeHLLC.cpp",
    "ref_fc_hyb": "src/flow/flow_models/four-eqn_conservative/Riemann_solvers/FlowModelRiemannSolverFourEqnConservativeHLLC-HLL.cpp",
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


def if_else_block(text: str, pattern: str, occurrence: int = -1) -> str:
    """The `if (...) {...} else {...}` of `text` whose `if` starts at the given occurrence of regex `pattern`."""
    m = list(re.finditer(pattern, text))[occurrence]

    def close_of(open_at: int) -> int:
        depth = 0
        for i in range(open_at, len(text)):
            if text[i] == "{":
                depth += 1
            elif text[i] == "}":
                depth -= 1
                if depth == 0:
                    return i
        raise ValueError("unbalanced braces")
    end_if = close_of(text.index("{", m.end()))
    rest = text[end_if + 1:]
    assert rest.lstrip().startswith("else"), rest[:40]
    end_else = close_of(text.index("{", end_if + 1))
    return text[m.start():end_else + 1]


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


def path_statements2() -> str:
    """Second group, same technique: velocity and specific internal energy (FlowModelSingleSpecies.cpp:2824-2826,
    3049-3051), face averages, characteristic projection and its inverse of the single-species 3-D x direction
    (FlowModelBasicUtilitiesSingleSpecies.cpp:5000-5001, 6324-6329, 7373-7379), RK update (Euler.cpp:1479, 1544-1548)."""
    with open(os.path.join(REF, "src/flow/flow_models/single-species/FlowModelSingleSpecies.cpp")) as fh:
        fm = fh.read()
    with open(os.path.join(REF, "src/flow/flow_models/single-species/FlowModelBasicUtilitiesSingleSpecies.cpp")) as fh:
        bu = fh.read()
    with open(os.path.join(REF, "src/apps/Euler/Euler.cpp")) as fh:
        eu = fh.read()
    s_u = statement(fm, r"u\[idx_velocity\] = rho_u\[idx\]/rho\[idx\]")
    s_v = statement(fm, r"v\[idx_velocity\] = rho_v\[idx\]/rho\[idx\]")
    s_w = statement(fm, r"w\[idx_velocity\] = rho_w\[idx\]/rho\[idx\]")
    s_e = statement(fm, r"epsilon\[idx_internal_energy\] = E\[idx\]/rho\[idx\] -\s*double\(1\)/double\(2\)\*\(u\[idx_velocity\]\*u\[idx_velocity\] \+ v\[idx_velocity\]\*v\[idx_velocity\] \+")
    s_ra = statement(bu, r"rho_average\[idx_face_x\] = double\(1\)/double\(2\)\*\(rho\[idx_L\] \+ rho\[idx_R\]\)")
    s_ca = statement(bu, r"c_average\[idx_face_x\] = double\(1\)/double\(2\)\*\(c\[idx_sound_speed_L\] \+ c\[idx_sound_speed_R\]\)")
    proj = [statement(bu, r"W\[0\]\[idx_face\] = -double\(1\)/double\(2\)\*rho_average\[idx_face\]\*c_average\[idx_face\]\*V\[1\]\[idx_vel\]"),
            statement(bu, r"W\[1\]\[idx_face\] = V\[0\]\[idx_rho\] - double\(1\)/\(c_average"),
            statement(bu, r"W\[2\]\[idx_face\] = V\[2\]\[idx_vel\]"),
            statement(bu, r"W\[3\]\[idx_face\] = V\[3\]\[idx_vel\]"),
            statement(bu, r"W\[4\]\[idx_face\] = double\(1\)/double\(2\)\*rho_average\[idx_face\]\*c_average\[idx_face\]\*V\[1\]\[idx_vel\]")]
    back = [statement(bu, r"V\[0\]\[idx_face\] = double\(1\)/\(c_average\[idx_face\]\*c_average\[idx_face\]\)\*W\[0\]\[idx_face\]"),
            statement(bu, r"V\[1\]\[idx_face\] = -double\(1\)/\(rho_average\[idx_face\]\*c_average\[idx_face\]\)\*W\[0\]\[idx_face\]"),
            statement(bu, r"V\[2\]\[idx_face\] = W\[2\]\[idx_face\]"),
            statement(bu, r"V\[3\]\[idx_face\] = W\[3\]\[idx_face\]"),
            statement(bu, r"V\[4\]\[idx_face\] = W\[0\]\[idx_face\] \+ W\[4\]\[idx_face\]")]
    s_al = statement(eu, r"Q\[ei\]\[idx\] \+= alpha\[n\]\*Q_intermediate\[ei\]\[idx_intermediate\]")
    s_be = statement(eu, r"Q\[ei\]\[idx\] \+= beta\[n\]\*\s*\(-\(F_x_intermediate\[idx_flux_x_R\] - F_x_intermediate\[idx_flux_x_L\]\)/dx_0 -\s*\(F_y_intermediate\[idx_flux_y_T\] - F_y_intermediate\[idx_flux_y_B\]\)/dx_1 -\s*\(F_z")
    nl = "\n        "
    return f"""
extern "C" void ref_path_points2(const double in[32], double out[20])
{{
    {{
        const int idx = 0, idx_velocity = 0, idx_internal_energy = 0;
        const double rho[1] = {{in[0]}}, rho_u[1] = {{in[1]}}, rho_v[1] = {{in[2]}}, rho_w[1] = {{in[3]}}, E[1] = {{in[4]}};
        double u[1], v[1], w[1], epsilon[1];
        {s_u}
        {s_v}
        {s_w}
        {s_e}
        out[0] = u[0]; out[1] = v[0]; out[2] = w[0]; out[3] = epsilon[0];
    }}
    double rho_average[1], c_average[1];
    {{
        const int idx_face_x = 0, idx_L = 0, idx_R = 1, idx_sound_speed_L = 0, idx_sound_speed_R = 1;
        const double rho[2] = {{in[5], in[6]}}, c[2] = {{in[7], in[8]}};
        {s_ra}
        {s_ca}
        out[4] = rho_average[0]; out[5] = c_average[0];
    }}
    {{
        const int idx_face = 0, idx_rho = 0, idx_vel = 0, idx_p = 0;
        double v0[1] = {{in[9]}}, v1[1] = {{in[10]}}, v2[1] = {{in[11]}}, v3[1] = {{in[12]}}, v4[1] = {{in[13]}};
        double* V[5] = {{v0, v1, v2, v3, v4}};
        double w0[1], w1[1], w2[1], w3[1], w4[1];
        double* W[5] = {{w0, w1, w2, w3, w4}};
        {nl.join(proj)}
        for (int e = 0; e < 5; e++) out[6 + e] = W[e][0];
    }}
    {{
        const int idx_face = 0;
        double w0[1] = {{in[14]}}, w1[1] = {{in[15]}}, w2[1] = {{in[16]}}, w3[1] = {{in[17]}}, w4[1] = {{in[18]}};
        double* W[5] = {{w0, w1, w2, w3, w4}};
        double v0[1], v1[1], v2[1], v3[1], v4[1];
        double* V[5] = {{v0, v1, v2, v3, v4}};
        {nl.join(back)}
        for (int e = 0; e < 5; e++) out[11 + e] = V[e][0];
    }}
    {{
        const int ei = 0, n = 0, idx = 0, idx_intermediate = 0, idx_source = 0;
        const int idx_flux_x_R = 0, idx_flux_x_L = 1, idx_flux_y_T = 0, idx_flux_y_B = 1, idx_flux_z_F = 0, idx_flux_z_B = 1;
        double q0[1] = {{in[19]}};
        double* Q[1] = {{q0}};
        const double alpha[1] = {{in[20]}}, beta[1] = {{in[22]}};
        double qi[1] = {{in[21]}};
        double* Q_intermediate[1] = {{qi}};
        const double F_x_intermediate[2] = {{in[23], in[24]}}, F_y_intermediate[2] = {{in[25], in[26]}}, F_z_intermediate[2] = {{in[27], in[28]}};
        const double dx_0 = in[29], dx_1 = in[30], dx_2 = in[31];
        const double S_intermediate[1] = {{in[18]}};
        {s_al}
        out[16] = Q[0][0];
        {s_be}
        out[17] = Q[0][0];
        out[18] = 0.0; out[19] = 0.0;
    }}
}}
"""


def line_range(text: str, first: int, last: int) -> str:
    """Lines first..last (1-based, inclusive) of text: restricts statement() to one branch of a dimension switch."""
    return "\n".join(text.split("\n")[first - 1:last])


def path_statements3() -> str:
    """Third group, five-eqn Allaire with two species, 3-D, x direction: mixture density
    (EquationOfStateMixingRules.cpp:871), mass fractions, velocity, internal energy (FlowModelFiveEqnAllaire.cpp:3965,
    4188-4190, 4428-4430), mixture gamma from all stored volume fractions (EquationOfStateMixingRulesIdealGas.cpp:7544,
    7565, 7586), pressure, Gruneisen parameter, Psi (EquationOfStateIdealGas.cpp:5756, 8157, 8308), sound speed
    (FlowModelFiveEqnAllaire.cpp:4700-4853), face averages, projection and back-projection
    (FlowModelBasicUtilitiesFiveEqnAllaire.cpp:7952-7996, 8848-8921, 9700-9760), advective source
    (ConvectiveFluxReconstructorWCNS56-HLLC-HLL.cpp:2623-2641)."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    fm = rd("src/flow/flow_models/five-eqn_Allaire/FlowModelFiveEqnAllaire.cpp")
    bu = rd("src/flow/flow_models/five-eqn_Allaire/FlowModelBasicUtilitiesFiveEqnAllaire.cpp")
    mr = rd("src/util/mixing_rules/equations_of_state/EquationOfStateMixingRules.cpp")
    mi = line_range(rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateMixingRulesIdealGas.cpp"), 7515, 7590)
    ig = rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateIdealGas.cpp")
    w56 = rd(SOURCES["ref_weno"].replace("WCNS5-JS-HLLC-HLL", "WCNS56-HLLC-HLL"))
    s_rho = statement(mr, r"rho\[idx_mixture_density\] \+= Z_rho\[si\]\[idx_partial_densities\]")
    s_Y = statement(fm, r"Y\[si\]\[idx_mass_fractions\] = Z_rho\[si\]\[idx\]/rho\[idx_density\]")
    s_u = statement(fm, r"u\[idx_velocity\] = rho_u\[idx\]/rho\[idx_density\]")
    s_v = statement(fm, r"v\[idx_velocity\] = rho_v\[idx\]/rho\[idx_density\]")
    s_w = statement(fm, r"w\[idx_velocity\] = rho_w\[idx\]/rho\[idx_density\]")
    s_e = statement(fm, r"epsilon\[idx_internal_energy\] = E\[idx\]/rho\[idx_density\] -\s*double\(1\)/double\(2\)\*\(u\[idx_velocity\]\*u\[idx_velocity\] \+ v\[idx_velocity\]\*v\[idx_velocity\] \+")
    s_ood = statement(mi, r"const double one_over_denominator = double\(1\)/\(d_species_gamma\[si\] - double\(1\)\)")
    s_xi = statement(mi, r"gamma\[idx_mixture_thermo_properties\] \+= Z\[si\]\[idx_volume_fractions\]\*one_over_denominator")
    s_gm = statement(mi, r"gamma\[idx_mixture_thermo_properties\] = double\(1\)/gamma\[idx_mixture_thermo_properties\] \+ double\(1\)")
    s_p = statement(ig, r"p\[idx_pressure\] = \(gamma\[idx_thermo_properties\] - double\(1\)\)\*rho\[idx_density\]\*\s*epsilon\[idx_internal_energy\]")
    s_Gr = statement(ig, r"Gamma\[idx_gruneisen_parameter\] = gamma\[idx_thermo_properties\] - double\(1\)")
    s_Psi = statement(ig, r"Psi\[idx_partial_pressure_partial_density\] = p\[idx_pressure\]/rho\[idx_density\]")
    s_c0 = statement(fm, r"c\[idx_sound_speed\] = Gamma\[idx_sound_speed\]\*p\[idx_pressure\]/rho\[idx_density\]")
    s_c1 = statement(fm, r"c\[idx_sound_speed\] \+= Y\[si\]\[idx_mass_fractions\]\*Psi\[si\]\[idx_sound_speed\]")
    s_c2 = statement(fm, r"c\[idx_sound_speed\] = sqrt\(c\[idx_sound_speed\]\)")
    s_za = statement(bu, r"Z_rho_average\[si\]\[idx_face_x\] = double\(1\)/double\(2\)\*\(Z_rho\[si\]\[idx_L\] \+ Z_rho\[si\]\[idx_R\]\)")
    s_ra = statement(bu, r"rho_average\[idx_face_x\] = double\(1\)/double\(2\)\*\(rho\[idx_density_L\] \+ rho\[idx_density_R\]\)")
    s_ca = statement(bu, r"c_average\[idx_face_x\] = double\(1\)/double\(2\)\*\(c\[idx_sound_speed_L\] \+ c\[idx_sound_speed_R\]\)")
    px = line_range(bu, 8800, 8925)       # 3-D, x direction of the projection
    s_w0 = statement(px, r"W\[0\]\[idx_face\] = V\[d_num_species\]\[idx_vel\] -")
    s_wsi = statement(px, r"W\[1 \+ si\]\[idx_face\] = V\[si\]\[idx_Z_rho\] - Z_rho_average")
    s_wv = statement(px, r"W\[d_num_species \+ 1\]\[idx_face\] = V\[d_num_species \+ 1\]\[idx_vel\]")
    s_ww = statement(px, r"W\[d_num_species \+ 2\]\[idx_face\] = V\[d_num_species \+ 2\]\[idx_vel\]")
    s_wz = statement(px, r"W\[d_num_species \+ 3 \+ si\]\[idx_face\] = V\[d_num_species \+ 4 \+ si\]\[idx_Z\]")
    s_wl = statement(px, r"W\[2\*d_num_species \+ 2\]\[idx_face\] = V\[d_num_species\]\[idx_vel\] \+")
    bx = line_range(bu, 9660, 9765)       # 3-D, x direction of the back-projection
    s_vsi = statement(bx, r"V\[si\]\[idx_face\] = -double\(1\)/double\(2\)\*Z_rho_average")
    s_vz = statement(bx, r"V\[d_num_species \+ 4 \+ si\]\[idx_face\] = W\[d_num_species \+ 3 \+ si\]\[idx_face\]")
    s_vu = statement(bx, r"V\[d_num_species\]\[idx_face\] = double\(1\)/double\(2\)\*W\[0\]\[idx_face\] \+")
    s_vv = statement(bx, r"V\[d_num_species \+ 1\]\[idx_face\] = W\[d_num_species \+ 1\]\[idx_face\]")
    s_vw = statement(bx, r"V\[d_num_species \+ 2\]\[idx_face\] = W\[d_num_species \+ 2\]\[idx_face\]")
    s_vp = statement(bx, r"V\[d_num_species \+ 3\]\[idx_face\] = -double\(1\)/double\(2\)\*rho_average")
    s_S = statement(w56, r"S\[idx_cell_nghost\] \+= dt\*Q\[ei\]\[idx_cell_wghost\]\*\(")
    return f"""
extern "C" void ref_path_points3(const double in[56], double out[32])
{{
    const int d_num_species = 2, d_num_eqn = 7;
    {{
        const int idx = 0, idx_density = 0, idx_mixture_density = 0, idx_partial_densities = 0, idx_mass_fractions = 0;
        const int idx_velocity = 0, idx_internal_energy = 0, idx_mixture_thermo_properties = 0, idx_volume_fractions = 0;
        const int idx_thermo_properties = 0, idx_pressure = 0, idx_gruneisen_parameter = 0, idx_sound_speed = 0;
        const int idx_partial_pressure_partial_density = 0;
        double zr0[1] = {{in[0]}}, zr1[1] = {{in[1]}}, z0[1] = {{in[6]}}, z1[1] = {{in[7]}};
        double* Z_rho[2] = {{zr0, zr1}};
        double* Z[2] = {{z0, z1}};
        const double rho_u[1] = {{in[2]}}, rho_v[1] = {{in[3]}}, rho_w[1] = {{in[4]}}, E[1] = {{in[5]}};
        const double d_species_gamma[2] = {{in[8], in[9]}};
        double rho[1] = {{0.0}}, y0[1], y1[1], u[1], v[1], w[1], epsilon[1], gamma[1] = {{0.0}}, p[1], Gamma[1], c[1];
        double psi0[1], psi1[1];
        double* Y[2] = {{y0, y1}};
        double* Psi[2] = {{psi0, psi1}};
        for (int si = 0; si < d_num_species; si++) {{ {s_rho} }}
        for (int si = 0; si < d_num_species; si++) {{ {s_Y} }}
        {s_u}
        {s_v}
        {s_w}
        {s_e}
        for (int si = 0; si < d_num_species; si++) {{
            {s_ood}
            {s_xi}
        }}
        {s_gm}
        {s_p}
        {s_Gr}
        for (int si = 0; si < d_num_species; si++) {{
            double* Psi_si = Psi[si];
            double* Psi = Psi_si;      /* the reference fills one species' array per call of the single-species EOS */
            {s_Psi}
        }}
        {s_c0}
        for (int si = 0; si < d_num_species; si++) {{ {s_c1} }}
        {s_c2}
        out[0] = rho[0]; out[1] = y0[0]; out[2] = y1[0]; out[3] = u[0]; out[4] = v[0]; out[5] = w[0]; out[6] = epsilon[0];
        out[7] = gamma[0]; out[8] = p[0]; out[9] = c[0];
    }}
    double zra0[1], zra1[1], rho_average[1], c_average[1];
    double* Z_rho_average[2] = {{zra0, zra1}};
    {{
        const int idx_face_x = 0, idx_L = 0, idx_R = 1, idx_density_L = 0, idx_density_R = 1;
        const int idx_sound_speed_L = 0, idx_sound_speed_R = 1;
        const double zr0[2] = {{in[10], in[11]}}, zr1[2] = {{in[12], in[13]}}, rho[2] = {{in[14], in[15]}}, c[2] = {{in[16], in[17]}};
        const double* Z_rho[2] = {{zr0, zr1}};
        for (int si = 0; si < d_num_species; si++) {{ {s_za} }}
        {s_ra}
        {s_ca}
        out[10] = zra0[0]; out[11] = zra1[0]; out[12] = rho_average[0]; out[13] = c_average[0];
    }}
    {{
        const int idx_face = 0, idx_Z_rho = 0, idx_vel = 0, idx_p = 0, idx_Z = 0;
        double v_[7][1], w_[7][1];
        double *V[7], *W[7];
        for (int e = 0; e < 7; e++) {{ v_[e][0] = in[18 + e]; V[e] = v_[e]; W[e] = w_[e]; }}
        for (int si = 0; si < d_num_species; si++) {{ {s_wsi} }}
        for (int si = 0; si < d_num_species - 1; si++) {{ {s_wz} }}
        {s_w0}
        {s_wv}
        {s_ww}
        {s_wl}
        /* reference order of the characteristic variables: u-c/(rho c), partial densities, tangential velocities,
           volume fractions, u+c/(rho c); ours swaps nothing: out[14..20] = W[0..6] */
        for (int e = 0; e < 7; e++) out[14 + e] = W[e][0];
    }}
    {{
        const int idx_face = 0;
        double v_[7][1], w_[7][1];
        double *V[7], *W[7];
        for (int e = 0; e < 7; e++) {{ w_[e][0] = in[25 + e]; V[e] = v_[e]; W[e] = w_[e]; }}
        for (int si = 0; si < d_num_species; si++) {{ {s_vsi} }}
        for (int si = 0; si < d_num_species - 1; si++) {{ {s_vz} }}
        {s_vu}
        {s_vv}
        {s_vw}
        {s_vp}
        for (int e = 0; e < 7; e++) out[21 + e] = V[e][0];
    }}
    {{
        const int ei = 0, idx_cell_nghost = 0, idx_cell_wghost = 0;
        const int idx_midpoint_x_R = 0, idx_midpoint_x_L = 1, idx_midpoint_x_RR = 2, idx_midpoint_x_LL = 3;
        const int idx_cell_wghost_x_R = 0, idx_cell_wghost_x_L = 1;
        const int idx_midpoint_y_T = 0, idx_midpoint_y_B = 1, idx_midpoint_y_TT = 2, idx_midpoint_y_BB = 3;
        const int idx_cell_wghost_y_T = 0, idx_cell_wghost_y_B = 1;
        const int idx_midpoint_z_F = 0, idx_midpoint_z_B = 1, idx_midpoint_z_FF = 2, idx_midpoint_z_BB = 3;
        const int idx_cell_wghost_z_F = 0, idx_cell_wghost_z_B = 1;
        double S[1] = {{in[32]}};
        const double dt = in[33];
        double q0[1] = {{in[34]}};
        double* Q[1] = {{q0}};
        const double u_midpoint_x[4] = {{in[35], in[36], in[37], in[38]}}, u[2] = {{in[39], in[40]}};
        const double v_midpoint_y[4] = {{in[41], in[42], in[43], in[44]}}, v[2] = {{in[45], in[46]}};
        const double w_midpoint_z[4] = {{in[47], in[48], in[49], in[50]}}, w[2] = {{in[51], in[52]}};
        const double dx[3] = {{in[53], in[54], in[55]}};
        {s_S}
        out[28] = S[0]; out[29] = 0.0; out[30] = 0.0; out[31] = 0.0;
    }}
}}
"""


def path_statements4() -> str:
    """Fourth group: the thermodynamic state the five-eqn Riemann solver rebuilds from one interpolated side
    (FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp:5709-5941: rho, Y, c) with the mixture gamma from the ns-1 interpolated
    volume fractions (EquationOfStateMixingRulesIdealGas.cpp:7770-7829, 3-D), Gruneisen parameter, Psi and
    epsilon(p) (EquationOfStateIdealGas.cpp:8157, 8308, 6414)."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    rs = line_range(rd("src/flow/flow_models/five-eqn_Allaire/Riemann_solvers/FlowModelRiemannSolverFiveEqnAllaireHLLC.cpp"), 5640, 5990)
    mi = line_range(rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateMixingRulesIdealGas.cpp"), 7736, 7835)
    ig = rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateIdealGas.cpp")
    s_r0 = statement(rs, r"rho_x_L\[idx\] = double\(0\)")
    s_r1 = statement(rs, r"rho_x_L\[idx\] \+= V_x_L\[si\]\[idx\]")
    s_Y = statement(rs, r"Y_x_L\[si\]\[idx\] = V_x_L\[si\]\[idx\]/rho_x_L\[idx\]")
    s_Z = statement(rs, r"Z_x_L\[si\]\[idx\] = V_x_L\[d_num_species \+ 4 \+ si\]\[idx\]")
    s_ood = statement(mi, r"const double one_over_denominator = double\(1\)/\(d_species_gamma\[si\] - double\(1\)\)")
    s_xi = statement(mi, r"gamma\[idx_mixture_thermo_properties\] \+= Z\[si\]\[idx_volume_fractions\]\*one_over_denominator")
    s_zl = statement(mi, r"Z_last\[idx_volume_fractions_last\] -= Z\[si\]\[idx_volume_fractions\]")
    s_xl = statement(mi, r"gamma\[idx_mixture_thermo_properties\] \+= Z_last\[idx_volume_fractions_last\]/")
    s_gm = statement(mi, r"gamma\[idx_mixture_thermo_properties\] = double\(1\)/gamma\[idx_mixture_thermo_properties\] \+ double\(1\)")
    s_Gr = statement(ig, r"Gamma\[idx_gruneisen_parameter\] = gamma\[idx_thermo_properties\] - double\(1\)")
    s_Psi = statement(ig, r"Psi\[idx_partial_pressure_partial_density\] = p\[idx_pressure\]/rho\[idx_density\]")
    s_eps = statement(ig, r"epsilon\[idx_internal_energy\] = p\[idx_pressure\]/\(\(gamma\[idx_thermo_properties\] - double\(1\)\)\*")
    s_c0 = statement(rs, r"c_x_L\[idx\] = Gamma_x_L\[idx\]\*V_x_L\[d_num_species \+ 3\]\[idx\]/rho_x_L\[idx\]")
    s_c1 = statement(rs, r"c_x_L\[idx\] \+= Y_x_L\[si\]\[idx\]\*Psi_x_L\[si\]\[idx\]")
    s_c2 = statement(rs, r"c_x_L\[idx\] = sqrt\(c_x_L\[idx\]\)")
    return f"""
extern "C" void ref_path_points4(const double in[12], double out[4])
{{
    const int d_num_species = 2;
    const int idx = 0, idx_mixture_thermo_properties = 0, idx_volume_fractions = 0, idx_volume_fractions_last = 0;
    const int idx_thermo_properties = 0, idx_gruneisen_parameter = 0, idx_partial_pressure_partial_density = 0;
    const int idx_pressure = 0, idx_density = 0, idx_internal_energy = 0;
    double v_[7][1];
    double* V_x_L[7];
    for (int e = 0; e < 7; e++) {{ v_[e][0] = in[e]; V_x_L[e] = v_[e]; }}
    const std::vector<double> d_species_gamma = {{in[7], in[8]}};
    double rho_x_L[1], y0[1], y1[1], z0[1], c_x_L[1], Gamma_x_L[1], psi0[1], psi1[1], epsilon[1];
    double* Y_x_L[2] = {{y0, y1}};
    double* Z_x_L[1] = {{z0}};
    double* Psi_x_L[2] = {{psi0, psi1}};
    {s_r0}
    for (int si = 0; si < d_num_species; si++) {{ {s_r1} }}
    for (int si = 0; si < d_num_species; si++) {{ {s_Y} }}
    for (int si = 0; si < d_num_species - 1; si++) {{ {s_Z} }}
    {{
        /* EquationOfStateMixingRulesIdealGas: gamma starts at 0 and Z_last at 1 (fillAll), then */
        double gamma[1] = {{0.0}}, Z_last[1] = {{1.0}};
        double** Z = Z_x_L;
        for (int si = 0; si < d_num_species - 1; si++) {{
            {s_ood}
            {s_xi}
            {s_zl}
        }}
        {s_xl}
        {s_gm}
        double* Gamma = Gamma_x_L;
        {s_Gr}
        const double* p = V_x_L[d_num_species + 3];
        const double* rho = rho_x_L;
        for (int si = 0; si < d_num_species; si++) {{
            double* Psi = Psi_x_L[si];
            {s_Psi}
        }}
        {s_eps}
    }}
    {s_c0}
    for (int si = 0; si < d_num_species; si++) {{ {s_c1} }}
    {s_c2}
    out[0] = rho_x_L[0]; out[1] = c_x_L[0]; out[2] = epsilon[0]; out[3] = 0.0;
}}
"""


def path_statements5() -> str:
    """Fifth group: the bounds flag of one interpolated side.  Five-eqn, 3-D: the x block
    (FlowModelBasicUtilitiesFiveEqnAllaire.cpp:6400-6712) and the y block (:6713-7026; the z block :7027-7340 repeats
    the y block's statements) -- they differ in the species loop of the c^2 check, see side_bounded in hamers_oracle.c.
    Single-species, 3-D x (FlowModelBasicUtilitiesSingleSpecies.cpp:3311-3326).  The Gruneisen parameter and Psi come from
    the statements of the third group (all ns volume fractions)."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    bu = rd("src/flow/flow_models/five-eqn_Allaire/FlowModelBasicUtilitiesFiveEqnAllaire.cpp")
    ss = line_range(rd("src/flow/flow_models/single-species/FlowModelBasicUtilitiesSingleSpecies.cpp"), 3270, 3330)
    mi = line_range(rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateMixingRulesIdealGas.cpp"), 7515, 7590)
    ig = rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateIdealGas.cpp")
    s_ood = statement(mi, r"const double one_over_denominator = double\(1\)/\(d_species_gamma\[si\] - double\(1\)\)")
    s_xi = statement(mi, r"gamma\[idx_mixture_thermo_properties\] \+= Z\[si\]\[idx_volume_fractions\]\*one_over_denominator")
    s_gm = statement(mi, r"gamma\[idx_mixture_thermo_properties\] = double\(1\)/gamma\[idx_mixture_thermo_properties\] \+ double\(1\)")
    s_Gr = statement(ig, r"Gamma\[idx_gruneisen_parameter\] = gamma\[idx_thermo_properties\] - double\(1\)")
    s_Psi = statement(ig, r"Psi\[idx_partial_pressure_partial_density\] = p\[idx_pressure\]/rho\[idx_density\]")

    def fe_block(first: int, last: int, name: str) -> str:
        t = line_range(bu, first, last)
        s_z = statement(t, r"Z\[si\]\[idx_face\] = V\[d_num_species \+ d_dim\.getValue\(\) \+ 1 \+ si\]\[idx_face\]")
        s_zl = statement(t, r"Z\[d_num_species - 1\]\[idx_face\] -= Z\[si\]\[idx_face\]")
        b_z = if_else_block(t, r"if \(Z\[si\]\[idx_face\] > d_Z_bound_lo")
        b_zl = if_else_block(t, r"if \(Z\[d_num_species - 1\]\[idx_face\] > d_Z_bound_lo")
        s_rho = statement(t, r"rho\[idx_face\] \+= V\[si\]\[idx_face\]")
        s_Y = statement(t, r"Y\[si\]\[idx_face\] = V\[si\]\[idx_face\]/rho\[idx_face\]")
        b_Y = if_else_block(t, r"if \(Y\[si\]\[idx_face\] > d_Y_bound_lo")
        b_V = if_else_block(t, r"if \(V\[si\]\[idx_face\] > double\(0\)\)")
        s_p = statement(t, r"p\[idx_face\] = V\[d_num_species \+ d_dim\.getValue\(\)\]\[idx_face\]")
        s_c0 = statement(t, r"c_sq\[idx_face\] = Gamma\[idx_face\]\*p\[idx_face\]/rho\[idx_face\]", 0)
        # the statement inside the species loop: `+= Y Psi` in the x block, a second `= Gamma p / rho` in the y / z blocks
        s_c1 = statement(t, r"c_sq\[idx_face\] (\+= Y\[si\]\[idx_face\]\*Psi\[si\]\[idx_face\]|= Gamma\[idx_face\]\*p\[idx_face\]/rho\[idx_face\])", -1)
        b_c = if_else_block(t, r"if \(c_sq\[idx_face\] > double\(0\)\)")
        return f"""
static int {name}(const double* Vin, const double* gam)
{{
    const int d_num_species = 2, idx_face = 0;
    struct {{ int getValue() const {{ return 3; }} }} d_dim;
    const int idx_mixture_thermo_properties = 0, idx_volume_fractions = 0, idx_thermo_properties = 0;
    const int idx_gruneisen_parameter = 0, idx_partial_pressure_partial_density = 0, idx_pressure = 0, idx_density = 0;
    const double d_Z_bound_lo = REF_Z_BOUND_LO, d_Z_bound_up = REF_Z_BOUND_UP;
    const double d_Y_bound_lo = REF_Y_BOUND_LO, d_Y_bound_up = REF_Y_BOUND_UP;
    const std::vector<double> d_species_gamma = {{gam[0], gam[1]}};
    int are_bounded[1] = {{1}};                              /* bounded_flag->fillAll(1) */
    double v_[7][1], z_[2][1] = {{{{0.0}}, {{1.0}}}}, y_[2][1], psi_[2][1];     /* last volume fraction filled with 1 */
    double *V[7], *Z[2] = {{z_[0], z_[1]}}, *Y[2] = {{y_[0], y_[1]}}, *Psi[2] = {{psi_[0], psi_[1]}};
    for (int e = 0; e < 7; e++) {{ v_[e][0] = Vin[e]; V[e] = v_[e]; }}
    double rho[1] = {{0.0}}, p[1], Gamma[1], c_sq[1], gamma[1] = {{0.0}};
    for (int si = 0; si < d_num_species - 1; si++) {{
        {s_z}
        {s_zl}
        {b_z}
    }}
    {b_zl}
    for (int si = 0; si < d_num_species; si++) {{ {s_rho} }}
    for (int si = 0; si < d_num_species; si++) {{
        {s_Y}
        {b_Y}
    }}
    for (int si = 0; si < d_num_species; si++) {{ {b_V} }}
    {s_p}
    for (int si = 0; si < d_num_species; si++) {{
        {s_ood}
        {s_xi}
    }}
    {s_gm}
    {s_Gr}
    for (int si = 0; si < d_num_species; si++) {{
        double* Psi_si = Psi[si];
        double* Psi = Psi_si;
        {s_Psi}
    }}
    {s_c0}
    for (int si = 0; si < d_num_species; si++) {{ {s_c1} }}
    {b_c}
    return are_bounded[0];
}}
"""
    b_r = if_else_block(ss, r"if \(V\[0\]\[idx_face\] > double\(0\)\)", 0)
    b_pp = if_else_block(ss, r"if \(V\[d_num_eqn - 1\]\[idx_face\] > double\(0\)\)", 0)
    fe_hpp = rd("include/flow/flow_models/five-eqn_Allaire/FlowModelBasicUtilitiesFiveEqnAllaire.hpp")

    def bound(nm):
        return re.search(nm + r" = double\(([-0-9.]+)\)", fe_hpp).group(1)
    return f"""
#define REF_Y_BOUND_LO ({bound("d_Y_bound_lo")})
#define REF_Y_BOUND_UP ({bound("d_Y_bound_up")})
#define REF_Z_BOUND_LO ({bound("d_Z_bound_lo")})
#define REF_Z_BOUND_UP ({bound("d_Z_bound_up")})
{fe_block(6400, 6712, "ref_fe_bounded_x")}
{fe_block(6713, 7026, "ref_fe_bounded_y")}
{fe_block(7027, 7340, "ref_fe_bounded_z")}
extern "C" void ref_path_points5(const double in[16], double out[2])
{{
    const int dir = (int)in[9];
    out[0] = (double)(dir == 0 ? ref_fe_bounded_x(in, in + 7) : (dir == 1 ? ref_fe_bounded_y(in, in + 7) : ref_fe_bounded_z(in, in + 7)));
    {{
        const int d_num_eqn = 5, idx_face = 0;
        int are_bounded[1] = {{1}};
        double v_[5][1];
        double* V[5];
        for (int e = 0; e < 5; e++) {{ v_[e][0] = in[10 + e]; V[e] = v_[e]; }}
        {b_r}
        {b_pp}
        out[1] = (double)are_bounded[0];
    }}
}}
"""


def path_statements6() -> str:
    """Sixth group: max wave speeds (FlowModelSingleSpecies.cpp:4064, 4237, 4365) and the spectral radii / stable dt of
    Euler::computeSpectralRadiusesAndStableDtOnPatch (Euler.cpp:846-861, 3-D)."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    fm = rd("src/flow/flow_models/single-species/FlowModelSingleSpecies.cpp")
    eu = line_range(rd("src/apps/Euler/Euler.cpp"), 760, 893)
    s_lx = statement(fm, r"lambda_max_x\[idx_max_wave_speed_x\] = fabs\(u\[idx_velocity\]\) \+ c\[idx_sound_speed\]")
    s_ly = statement(fm, r"lambda_max_y\[idx_max_wave_speed_y\] = fabs\(v\[idx_velocity\]\) \+ c\[idx_sound_speed\]")
    s_lz = statement(fm, r"lambda_max_z\[idx_max_wave_speed_z\] = fabs\(w\[idx_velocity\]\) \+ c\[idx_sound_speed\]")
    s_rx = statement(eu, r"const double spectral_radius_x = max_lambda_x\[idx\]/dx_0")
    s_ry = statement(eu, r"const double spectral_radius_y = max_lambda_y\[idx\]/dx_1")
    s_rz = statement(eu, r"const double spectral_radius_z = max_lambda_z\[idx\]/dx_2")
    s_m3 = statement(eu, r"spectral_radiuses_and_dt_3 = fmax\(spectral_radiuses_and_dt_3,")
    s_dt = statement(eu, r"spectral_radiuses_and_dt\[3\] = double\(1\)/spectral_radiuses_and_dt\[3\]")
    return f"""
extern "C" void ref_path_points6(const double in[8], double out[6])
{{
    const int idx = 0, idx_velocity = 0, idx_sound_speed = 0;
    const int idx_max_wave_speed_x = 0, idx_max_wave_speed_y = 0, idx_max_wave_speed_z = 0;
    const double u[1] = {{in[0]}}, v[1] = {{in[1]}}, w[1] = {{in[2]}}, c[1] = {{in[3]}};
    const double dx_0 = in[4], dx_1 = in[5], dx_2 = in[6];
    double lambda_max_x[1], lambda_max_y[1], lambda_max_z[1];
    {s_lx}
    {s_ly}
    {s_lz}
    const double *max_lambda_x = lambda_max_x, *max_lambda_y = lambda_max_y, *max_lambda_z = lambda_max_z;
    {s_rx}
    {s_ry}
    {s_rz}
    out[0] = spectral_radius_x; out[1] = spectral_radius_y; out[2] = spectral_radius_z;
    double spectral_radiuses_and_dt_3 = -1.0;
    {s_m3}
    out[3] = spectral_radiuses_and_dt_3;            /* the sum as the reference associates it */
    spectral_radiuses_and_dt_3 = in[7];
    {s_m3}
    double spectral_radiuses_and_dt[4] = {{0.0, 0.0, 0.0, spectral_radiuses_and_dt_3}};
    out[4] = spectral_radiuses_and_dt[3];
    {s_dt}
    out[5] = spectral_radiuses_and_dt[3];
}}
"""


def void_member_function(text: str, cls: str, name: str) -> str:
    """The definition `void\n<cls>::<name>(...) const { ... }` in text, as text."""
    m = re.search(r"^void\s*\n" + cls + "::" + name + r"\(", text, flags=re.M)
    brace = text.index("{", m.end())
    depth, i = 0, brace
    while True:
        if text[i] == "{":
            depth += 1
        elif text[i] == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return text[m.start():i + 1]


def diffusive_kernels() -> str:
    """the six kernels of DiffusiveFluxReconstructorNodeSixthOrder (computeFirstDerivativesIn{X,Y,Z},
    reconstructFlux{X,Y,Z}, DiffusiveFluxReconstructorNodeSixthOrder.cpp:65-939) compiled verbatim as members of a stub
    class (hier::IntVector / tbox::Dimension reduced to what the kernels use), driven over the same index ranges as the
    base class does (DiffusiveFluxReconstructorNode.cpp:1795-1812, 2276-2297); and the reference's statements for the
    temperature, the Prandtl conductivity and the diffusivities D_00..D_12 / D_00..D_09."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    cls = "DiffusiveFluxReconstructorNodeSixthOrder"
    src = rd("src/flow/diffusive_flux_reconstructors/node/DiffusiveFluxReconstructorNodeSixthOrder.cpp")
    names = ["computeFirstDerivativesInX", "computeFirstDerivativesInY", "computeFirstDerivativesInZ",
             "reconstructFluxX", "reconstructFluxY", "reconstructFluxZ"]
    bodies = "\n\n".join(void_member_function(src, cls, n) for n in names)
    iv = "const hier::IntVector&"
    der = f"(double*, const double* const, {iv}, {iv}, {iv}, {iv}, {iv}, {iv}, const double&) const;"
    rec = f"(double*, const double* const, {iv}, {iv}, {iv}, {iv}, {iv}, const double&) const;"
    decls = "\n".join(f"    void {n}{der if 'Derivatives' in n else rec}" for n in names)
    ig = rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateIdealGas.cpp")
    pr = rd("src/util/mixing_rules/equations_of_thermal_conductivity/Prandtl/EquationOfThermalConductivityPrandtl.cpp")
    du = rd("src/flow/flow_models/single-species/FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp")
    s_T = statement(ig, r"T\[idx_temperature\] = p\[idx_pressure\]/\(\(gamma - double\(1\)\)\*c_v\*rho\[idx_density\]\)")
    s_k = statement(pr, r"kappa\[idx_thermal_conductivity\] = c_p\*mu\[idx_min\]/Pr")
    du3 = line_range(du, 4255, 4285)
    du2 = line_range(du, 4175, 4198)
    D3 = "\n        ".join(statement(du3, r"D_%02d\[idx_diffusivities\] = " % m) for m in range(13))
    D2 = "\n        ".join(statement(du2, r"D_%02d\[idx_diffusivities\] = " % m) for m in range(10))
    return f"""
namespace ref_diff {{
namespace hier {{ struct IntVector {{ int v[3]; int operator[](int i) const {{ return v[i]; }} }}; }}
namespace tbox {{ struct Dimension {{ int d; explicit Dimension(int d_) : d(d_) {{}}
                  bool operator==(const Dimension& o) const {{ return d == o.d; }} }}; }}
struct {cls} {{
    tbox::Dimension d_dim;
    explicit {cls}(int dim) : d_dim(dim) {{}}
{decls}
}};
{bodies}
}}

extern "C" void ref_diff_derivative(int dim, int dir, const double* u, const int* n, double dx_inv, double* out)
{{
    using namespace ref_diff;
    const int g = 6;
    {cls} k(dim);
    hier::IntVector ng = {{{{g, g, dim == 3 ? g : 0}}}}, dims = {{{{n[0] + 2 * g, n[1] + 2 * g, dim == 3 ? n[2] + 2 * g : 1}}}};
    hier::IntVector lo = {{{{-g, -g, dim == 3 ? -g : 0}}}}, dd = dims;
    lo.v[dir] += 3;
    dd.v[dir] -= 6;
    if (dir == 0) k.computeFirstDerivativesInX(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
    if (dir == 1) k.computeFirstDerivativesInY(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
    if (dir == 2) k.computeFirstDerivativesInZ(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
}}

extern "C" void ref_diff_reconstruct(int dim, int dir, const double* F_node, const int* n, double dt, double* F_face)
{{
    using namespace ref_diff;
    const int g = 6;
    {cls} k(dim);
    hier::IntVector ng = {{{{g, g, dim == 3 ? g : 0}}}}, dims = {{{{n[0] + 2 * g, n[1] + 2 * g, dim == 3 ? n[2] + 2 * g : 1}}}};
    hier::IntVector lo = {{{{0, 0, 0}}}}, interior = {{{{n[0], n[1], dim == 3 ? n[2] : 1}}}}, dd = interior;
    dd.v[dir]++;
    if (dir == 0) k.reconstructFluxX(F_face, F_node, ng, dims, lo, dd, interior, dt);
    if (dir == 1) k.reconstructFluxY(F_face, F_node, ng, dims, lo, dd, interior, dt);
    if (dir == 2) k.reconstructFluxZ(F_face, F_node, ng, dims, lo, dd, interior, dt);
}}

extern "C" void ref_diff_point(int dim, const double in[10], double out[15])
{{
    /* in: gamma, c_v, rho, p, c_p, mu, Pr, mu_v | u, v, w(in[10] is not read in 2-D);  out: T, kappa, D[...] */
    const int idx_temperature = 0, idx_pressure = 0, idx_density = 0, idx_thermal_conductivity = 0, idx_min = 0;
    const int idx_diffusivities = 0, idx_shear_viscosity = 0, idx_bulk_viscosity = 0, idx_velocity = 0;
    const double gamma = in[0], c_v = in[1], rho[1] = {{in[2]}}, p[1] = {{in[3]}}, c_p = in[4], mu[1] = {{in[5]}}, Pr = in[6];
    const double mu_v[1] = {{in[7]}}, u[1] = {{in[8]}}, v[1] = {{in[9]}}, w[1] = {{dim == 3 ? in[10] : 0.0}};
    double T[1], kappa[1];
    {s_T}
    {s_k}
    out[0] = T[0]; out[1] = kappa[0];
    double D_00[1], D_01[1], D_02[1], D_03[1], D_04[1], D_05[1], D_06[1], D_07[1], D_08[1], D_09[1], D_10[1], D_11[1], D_12[1];
    if (dim == 3) {{
        {D3}
        const double* D[13] = {{D_00, D_01, D_02, D_03, D_04, D_05, D_06, D_07, D_08, D_09, D_10, D_11, D_12}};
        for (int m = 0; m < 13; m++) out[2 + m] = D[m][0];
    }} else {{
        {D2}
        const double* D[10] = {{D_00, D_01, D_02, D_03, D_04, D_05, D_06, D_07, D_08, D_09}};
        for (int m = 0; m < 10; m++) out[2 + m] = D[m][0];
        (void)w; (void)D_10; (void)D_11; (void)D_12;
    }}
}}
"""


def diffusive_term_tables() -> str:
    """FlowModelDiffusiveFluxUtilitiesSingleSpecies::getCellDataOfDiffusiveFluxVariablesForDerivative (:654-1503)
    and ::getCellDataOfDiffusiveFluxDiffusivities (:1505-2363) -- which derivative carries which diffusivity in which
    equation -- compiled verbatim as members of a stub class (FlowModel, pdat::CellData, tbox::Dimension, TBOX_ERROR reduced
    to what the two functions use)."""
    with open(os.path.join(REF, "src/flow/flow_models/single-species/FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp")) as fh:
        src = fh.read()
    cls = "FlowModelDiffusiveFluxUtilitiesSingleSpecies"
    f1 = void_member_function(src, cls, "getCellDataOfDiffusiveFluxVariablesForDerivative")
    f2 = void_member_function(src, cls, "getCellDataOfDiffusiveFluxDiffusivities")
    return f"""
#include <memory>
#include <string>
#include <sstream>
#include <vector>
#include <cstdlib>
namespace ref_diff_tables {{
#define TBOX_ERROR(X) do {{ std::ostringstream os_; os_ << X; std::abort(); }} while (0)
#define HAMERS_SHARED_PTR std::shared_ptr
namespace tbox {{ struct Dimension {{ int d; explicit Dimension(int d_) : d(d_) {{}}
                  bool operator==(const Dimension& o) const {{ return d == o.d; }} }}; }}
namespace pdat {{ template <class T> struct CellData {{ int tag; }}; }}
namespace DIRECTION {{ enum TYPE {{ X_DIRECTION = 0, Y_DIRECTION = 1, Z_DIRECTION = 2 }}; }}
struct FlowModel {{
    std::shared_ptr<pdat::CellData<double> > velocity, temperature;
    bool hasRegisteredPatch() const {{ return true; }}
    std::shared_ptr<pdat::CellData<double> > getCellData(const std::string& key) const
    {{
        return key == "VELOCITY" ? velocity : (key == "TEMPERATURE" ? temperature : std::shared_ptr<pdat::CellData<double> >());
    }}
}};
struct {cls} {{
    std::weak_ptr<FlowModel> d_flow_model;
    std::string d_object_name;
    tbox::Dimension d_dim;
    int d_num_eqn;
    bool d_cell_data_computed_diffusivities;
    std::shared_ptr<pdat::CellData<double> > d_data_diffusivities;
    explicit {cls}(int dim) : d_object_name("ref"), d_dim(dim), d_num_eqn(dim + 2), d_cell_data_computed_diffusivities(true) {{}}
    void getCellDataOfDiffusiveFluxVariablesForDerivative(std::vector<std::vector<std::shared_ptr<pdat::CellData<double> > > >&,
                                                          std::vector<std::vector<int> >&, const DIRECTION::TYPE&, const DIRECTION::TYPE&);
    void getCellDataOfDiffusiveFluxDiffusivities(std::vector<std::vector<std::shared_ptr<pdat::CellData<double> > > >&,
                                                 std::vector<std::vector<int> >&, const DIRECTION::TYPE&, const DIRECTION::TYPE&);
}};
{f1}

{f2}
}}

/* the terms of equation e of the node flux in direction fdir that carry a derivative in direction ddir: variable
 * (velocity component, or dim for the temperature) and index of the diffusivity; returns -1 if the two tables disagree */
extern "C" int ref_diff_terms(int dim, int fdir, int ddir, int e, int var[4], int diff[4])
{{
    using namespace ref_diff_tables;
    std::shared_ptr<FlowModel> fm(new FlowModel());
    fm->velocity.reset(new pdat::CellData<double>());
    fm->temperature.reset(new pdat::CellData<double>());
    {cls} u(dim);
    u.d_flow_model = fm;
    u.d_data_diffusivities.reset(new pdat::CellData<double>());
    std::vector<std::vector<std::shared_ptr<pdat::CellData<double> > > > vdata, ddata;
    std::vector<std::vector<int> > vidx, didx;
    u.getCellDataOfDiffusiveFluxVariablesForDerivative(vdata, vidx, (DIRECTION::TYPE)fdir, (DIRECTION::TYPE)ddir);
    u.getCellDataOfDiffusiveFluxDiffusivities(ddata, didx, (DIRECTION::TYPE)fdir, (DIRECTION::TYPE)ddir);
    if (vdata[e].size() != ddata[e].size() || vidx[e].size() != vdata[e].size() || didx[e].size() != ddata[e].size()) return -1;
    const int n = (int)vdata[e].size();
    for (int i = 0; i < n && i < 4; i++) {{
        if (ddata[e][i] != u.d_data_diffusivities) return -1;
        var[i] = vdata[e][i] == fm->velocity ? vidx[e][i] : (vdata[e][i] == fm->temperature && vidx[e][i] == 0 ? dim : -1);
        diff[i] = didx[e][i];
    }}
    return n;
}}
"""


def diffusive_dt_statements() -> str:
    """MAX_DIFFUSIVITY (FlowModelSingleSpecies.cpp:4661-4665) and the diffusive spectral radius / stable
    dt of NavierStokes::computeSpectralRadiusesAndStableDtOnPatch (NavierStokes.cpp:884-886 and 893 in 2-D, 1083-1086 and
    1089-1091 in 3-D): the reference's statements compiled verbatim."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    fm = rd("src/flow/flow_models/single-species/FlowModelSingleSpecies.cpp")
    ns = rd("src/apps/Navier-Stokes/NavierStokes.cpp")
    s_d1 = statement(fm, r"D_max\[idx_max_diffusivity\] = fmax\(mu\[idx_max_diffusivity\]/rho\[idx\],")
    s_d2 = statement(fm, r"D_max\[idx_max_diffusivity\] = fmax\(D_max\[idx_max_diffusivity\],")
    ns3, ns2 = line_range(ns, 1000, 1095), line_range(ns, 830, 900)
    s_r3 = statement(ns3, r"const double spectral_radius_diffusive = double\(2\)\*fmax\(")
    s_r2 = statement(ns2, r"const double spectral_radius_diffusive = double\(2\)\*fmax\(")
    s_m3 = statement(ns3, r"spectral_radiuses_and_dt\[3\] = fmax\(spectral_radius_tmp, spectral_radiuses_and_dt\[3\]\)")
    s_t3 = statement(ns3, r"spectral_radiuses_and_dt\[3\] = double\(1\)/\(spectral_radiuses_and_dt\[3\] \+ HAMERS_EPSILON\)")
    return f"""
extern "C" void ref_diff_dt_point(int dim, const double in[9], double out[3])
{{
    /* in: mu, mu_v, kappa, c_p, rho, dx_0, dx_1, dx_2, max sum of the acoustic radii;  out: D_max, diffusive radius, dt */
    const int idx = 0, idx_max_diffusivity = 0;
    const double mu[1] = {{in[0]}}, mu_v[1] = {{in[1]}}, kappa[1] = {{in[2]}}, c_p[1] = {{in[3]}}, rho[1] = {{in[4]}};
    const double dx_0 = in[5], dx_1 = in[6], dx_2 = in[7];
    double D_max[1];
    {s_d1}
    {s_d2}
    out[0] = D_max[0];
    const double* max_D = D_max;
    double radius;
    if (dim == 3) {{
        {s_r3}
        radius = spectral_radius_diffusive;
    }} else {{
        {s_r2}
        radius = spectral_radius_diffusive;
        (void)dx_2;
    }}
    out[1] = radius;
    double spectral_radius_tmp = radius;
    double spectral_radiuses_and_dt[4] = {{0.0, 0.0, 0.0, in[8]}};
    {s_m3}
    {s_t3}
    out[2] = spectral_radiuses_and_dt[3];
}}
"""


def midpoint_kernels() -> str:
    """SURVEY row f4, midpoint family: the twelve kernels of DiffusiveFluxReconstructorMidpointSixthOrder
    (computeFirstDerivativesIn{X,Y,Z}AtMidpoint{X,Y,Z}, computeFirstDerivativesIn{X,Y,Z}AtNode,
    interpolateDataFromNodeToMidpoint{X,Y,Z}, reconstructFlux{X,Y,Z};
    DiffusiveFluxReconstructorMidpointSixthOrder.cpp:68-1799) compiled verbatim as members of a stub class and driven over
    the whole range their stencils allow on a ghost box of width g; the reference's statements for the side diffusivities
    (FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp:2581-2589, 2615-2623, 2682-2691, 2725-2734, 2768-2777); and
    FlowModelDiffusiveFluxUtilitiesSingleSpecies::getSideDataOfDiffusiveFluxDiffusivities (:2799-3657: which side diffusivity
    multiplies which derivative in which equation) compiled verbatim as a member of a stub class."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    cls = "DiffusiveFluxReconstructorMidpointSixthOrder"
    src = rd("src/flow/diffusive_flux_reconstructors/midpoint/DiffusiveFluxReconstructorMidpointSixthOrder.cpp")
    XYZ = "XYZ"
    der_mid = [f"computeFirstDerivativesIn{c}AtMidpoint{c}" for c in XYZ]
    der_node = [f"computeFirstDerivativesIn{c}AtNode" for c in XYZ]
    interp = [f"interpolateDataFromNodeToMidpoint{c}" for c in XYZ]
    recon = [f"reconstructFlux{c}" for c in XYZ]
    names = der_mid + der_node + interp + recon
    bodies = "\n\n".join(void_member_function(src, cls, n) for n in names)
    iv = "const hier::IntVector&"
    sig_der = f"(double*, const double* const, {iv}, {iv}, {iv}, {iv}, {iv}, {iv}, const double&) const;"
    sig_int = f"(double*, const double* const, {iv}, {iv}, {iv}, {iv}, {iv}, {iv}) const;"
    sig_rec = f"(double*, const double* const, {iv}, {iv}, {iv}, {iv}, {iv}, const double&) const;"
    decls = "\n".join(f"    void {n}{sig_der if 'Derivatives' in n else (sig_int if 'interpolate' in n else sig_rec)}" for n in names)
    du = rd("src/flow/flow_models/single-species/FlowModelDiffusiveFluxUtilitiesSingleSpecies.cpp")

    def side_block(first, last, nd):
        blk = line_range(du, first, last)
        return "\n        ".join(statement(blk, r"D_%02d\[idx_diffusivities\] = " % m) for m in range(nd))
    S3 = [side_block(2680, 2692, 8), side_block(2723, 2735, 8), side_block(2766, 2778, 8)]
    S2 = [side_block(2579, 2590, 7), side_block(2613, 2624, 7)]
    fside = void_member_function(du, "FlowModelDiffusiveFluxUtilitiesSingleSpecies", "getSideDataOfDiffusiveFluxDiffusivities")
    return f"""
namespace ref_mid {{
namespace hier {{ struct IntVector {{ int v[3]; int operator[](int i) const {{ return v[i]; }} }}; }}
namespace tbox {{ struct Dimension {{ int d; explicit Dimension(int d_) : d(d_) {{}}
                  bool operator==(const Dimension& o) const {{ return d == o.d; }} }}; }}
struct {cls} {{
    tbox::Dimension d_dim;
    explicit {cls}(int dim) : d_dim(dim) {{}}
{decls}
}};
{bodies}
}}

/* kind 0: derivative along dir at the midpoints of dir (midpoints 3 - g .. n + g - 3 of dir, all cells of the ghost box in the
 * other directions); kind 1: derivative along dir at the nodes (nodes 3 - g .. n + g - 4 of dir); kind 2: interpolation
 * along dir from the nodes to the midpoints of dir.  Midpoint arrays have one more entry along dir than the ghost box. */
extern "C" void ref_mid_kernel(int kind, int dim, int dir, int g, const double* u, const int* n, double dx_inv, double* out)
{{
    using namespace ref_mid;
    {cls} k(dim);
    hier::IntVector ng = {{{{g, g, dim == 3 ? g : 0}}}}, dims = {{{{n[0] + 2 * g, n[1] + 2 * g, dim == 3 ? n[2] + 2 * g : 1}}}};
    hier::IntVector lo = {{{{-g, -g, dim == 3 ? -g : 0}}}}, dd = dims;
    lo.v[dir] += 3;
    dd.v[dir] -= (kind == 1) ? 6 : 5;
    if (kind == 0) {{
        if (dir == 0) k.computeFirstDerivativesInXAtMidpointX(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
        if (dir == 1) k.computeFirstDerivativesInYAtMidpointY(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
        if (dir == 2) k.computeFirstDerivativesInZAtMidpointZ(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
    }} else if (kind == 1) {{
        if (dir == 0) k.computeFirstDerivativesInXAtNode(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
        if (dir == 1) k.computeFirstDerivativesInYAtNode(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
        if (dir == 2) k.computeFirstDerivativesInZAtNode(out, u, ng, ng, dims, dims, lo, dd, dx_inv);
    }} else {{
        if (dir == 0) k.interpolateDataFromNodeToMidpointX(out, u, ng, ng, dims, dims, lo, dd);
        if (dir == 1) k.interpolateDataFromNodeToMidpointY(out, u, ng, ng, dims, dims, lo, dd);
        if (dir == 2) k.interpolateDataFromNodeToMidpointZ(out, u, ng, ng, dims, dims, lo, dd);
    }}
}}

/* faces 0 .. n of dir from the midpoint fluxes (a midpoint array on the ghost box of width g); F_face is "+=" on the caller's
 * zeros like the reference's fillAll(0) */
extern "C" void ref_mid_reconstruct(int dim, int dir, int g, const double* F_mid, const int* n, double dt, double* F_face)
{{
    using namespace ref_mid;
    {cls} k(dim);
    hier::IntVector ng = {{{{g, g, dim == 3 ? g : 0}}}}, dims = {{{{n[0] + 2 * g, n[1] + 2 * g, dim == 3 ? n[2] + 2 * g : 1}}}};
    hier::IntVector lo = {{{{0, 0, 0}}}}, interior = {{{{n[0], n[1], dim == 3 ? n[2] : 1}}}}, dd = interior;
    dd.v[dir]++;
    if (dir == 0) k.reconstructFluxX(F_face, F_mid, ng, dims, lo, dd, interior, dt);
    if (dir == 1) k.reconstructFluxY(F_face, F_mid, ng, dims, lo, dd, interior, dt);
    if (dir == 2) k.reconstructFluxZ(F_face, F_mid, ng, dims, lo, dd, interior, dt);
}}

/* side diffusivities of direction dir from the values interpolated to a midpoint; in: mu, mu_v, kappa, u, v, w */
extern "C" void ref_mid_side_diffusivities(int dim, int dir, const double in[6], double out[8])
{{
    const int idx_diffusivities = 0, idx_var_data = 0;
    double D_00[1], D_01[1], D_02[1], D_03[1], D_04[1], D_05[1], D_06[1], D_07[1] = {{0.0}};
    const double mu_x[1] = {{in[0]}}, mu_v_x[1] = {{in[1]}}, kappa_x[1] = {{in[2]}}, u_x[1] = {{in[3]}}, v_x[1] = {{in[4]}}, w_x[1] = {{in[5]}};
    const double *mu_y = mu_x, *mu_v_y = mu_v_x, *kappa_y = kappa_x, *u_y = u_x, *v_y = v_x, *w_y = w_x;
    const double *mu_z = mu_x, *mu_v_z = mu_v_x, *kappa_z = kappa_x, *u_z = u_x, *v_z = v_x, *w_z = w_x;
    (void)w_x; (void)w_y; (void)mu_z; (void)mu_v_z; (void)kappa_z; (void)u_z; (void)v_z; (void)w_z;
    if (dim == 3) {{
        if (dir == 0) {{ {S3[0]} }}
        if (dir == 1) {{ {S3[1]} }}
        if (dir == 2) {{ {S3[2]} }}
    }} else {{
        if (dir == 0) {{ {S2[0]} }}
        if (dir == 1) {{ {S2[1]} }}
    }}
    const double* D[8] = {{D_00, D_01, D_02, D_03, D_04, D_05, D_06, D_07}};
    for (int m = 0; m < (dim == 3 ? 8 : 7); m++) out[m] = D[m][0];
}}

namespace ref_diff_side_tables {{
#define HAMERS_SHARED_PTR std::shared_ptr
namespace tbox {{ struct Dimension {{ int d; explicit Dimension(int d_) : d(d_) {{}}
                  bool operator==(const Dimension& o) const {{ return d == o.d; }} }}; }}
namespace pdat {{ template <class T> struct SideData {{ int tag; }}; }}
namespace DIRECTION {{ enum TYPE {{ X_DIRECTION = 0, Y_DIRECTION = 1, Z_DIRECTION = 2 }}; }}
struct FlowModel {{ bool hasRegisteredPatch() const {{ return true; }} }};
struct FlowModelDiffusiveFluxUtilitiesSingleSpecies {{
    std::weak_ptr<FlowModel> d_flow_model;
    std::string d_object_name;
    tbox::Dimension d_dim;
    int d_num_eqn;
    bool d_side_data_diffusivities_computed;
    std::shared_ptr<pdat::SideData<double> > d_side_data_diffusivities;
    explicit FlowModelDiffusiveFluxUtilitiesSingleSpecies(int dim) : d_object_name("ref"), d_dim(dim), d_num_eqn(dim + 2), d_side_data_diffusivities_computed(true) {{}}
    void getSideDataOfDiffusiveFluxDiffusivities(std::vector<std::vector<std::shared_ptr<pdat::SideData<double> > > >&,
                                                 std::vector<std::vector<int> >&, const DIRECTION::TYPE&, const DIRECTION::TYPE&);
}};
{fside}
}}

/* indices of the side diffusivities that multiply the derivatives in direction ddir in equation e of the flux in direction
 * fdir, in the reference's order (the variables are those of ref_diff_terms: one function serves both reconstructor families) */
extern "C" int ref_diff_side_terms(int dim, int fdir, int ddir, int e, int diff[4])
{{
    using namespace ref_diff_side_tables;
    std::shared_ptr<FlowModel> fm(new FlowModel());
    FlowModelDiffusiveFluxUtilitiesSingleSpecies u(dim);
    u.d_flow_model = fm;
    u.d_side_data_diffusivities.reset(new pdat::SideData<double>());
    std::vector<std::vector<std::shared_ptr<pdat::SideData<double> > > > ddata;
    std::vector<std::vector<int> > didx;
    u.getSideDataOfDiffusiveFluxDiffusivities(ddata, didx, (DIRECTION::TYPE)fdir, (DIRECTION::TYPE)ddir);
    if (ddata[e].size() != didx[e].size()) return -1;
    const int n = (int)ddata[e].size();
    for (int i = 0; i < n && i < 4; i++) {{
        if (ddata[e][i] != u.d_side_data_diffusivities) return -1;
        diff[i] = didx[e][i];
    }}
    return n;
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


FC_WRAPPER = r"""
extern "C" int ref_riemann_point_fc(int dim, int ns, int dir, const double* V_L, const double* V_R,
                                    double rho_L, double rho_R, double c_L, double c_R, double eps_L, double eps_R,
                                    double* F_HLLC, double* F_HYB)
{
    /* FlowModelRiemannSolverFourEqnConservativeHLLC.cpp / ...HLLC-HLL.cpp point kernels on one face (idx = idx_flux = 0) */
    const int neq = dim + 1 + ns;
    double vl[16], vr[16], f1[16], f2[16];
    double *VL[16], *VR[16], *F1[16], *F2[16];
    for (int e = 0; e < neq; e++) { vl[e] = V_L[e]; vr[e] = V_R[e]; VL[e] = &vl[e]; VR[e] = &vr[e]; F1[e] = &f1[e]; F2[e] = &f2[e]; }
    double s_minus = 0, s_plus = 0, s_star = 0, Chi = 0;
    const int key = dim*10 + dir;
#define FC_CALL(D, N) \
    ref_fc_hllc::computeLocalConvectiveFluxIn##D##DirectionFromPrimitiveVariablesHLLC##N(F1, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0, ns, neq); \
    ref_fc_hyb::computeLocalConvectiveFluxIn##D##DirectionFromPrimitiveVariablesHLLC_HLL##N(F2, VL, VR, &rho_L, &rho_R, &c_L, &c_R, &eps_L, &eps_R, s_minus, s_plus, s_star, Chi, 0, 0, ns, neq);
    switch (key) {
    case 20: FC_CALL(X, 2D) break;
    case 21: FC_CALL(Y, 2D) break;
    case 30: FC_CALL(X, 3D) break;
    case 31: FC_CALL(Y, 3D) break;
    case 32: FC_CALL(Z, 3D) break;
    default: return -1;
    }
#undef FC_CALL
    for (int e = 0; e < neq; e++) { F_HLLC[e] = f1[e]; F_HYB[e] = f2[e]; }
    return 0;
}
"""


def path_statements7() -> str:
    """Seventh group, four-eqn conservative model with two species, 3-D: mixture density
    (EquationOfStateMixingRules.cpp), mass fractions, velocity, internal energy (FlowModelFourEqnConservative.cpp:4107,
    4767-4769, 5007), species c_p / c_v (EquationOfStateMixingRulesIdealGas.cpp:108-119), mixture c_p, c_v, gamma from the
    mass fractions (the array form, :7080-7120), pressure, epsilon from p, Gruneisen parameter (EquationOfStateIdealGas.cpp),
    Psi_i (EquationOfStateMixingRulesIdealGas.cpp:6540-6545), sound speed (FlowModelFourEqnConservative.cpp:5373-5421);
    the bounds flag of an interpolated side (FlowModelBasicUtilitiesFourEqnConservative.cpp:4390-4500, 3-D x block);
    projection / back-projection in x (FlowModelBasicUtilitiesFourEqnConservative.cpp:5896-5975, 6640-6700)."""
    def rd(rel):
        with open(os.path.join(REF, rel)) as fh:
            return fh.read()
    fm = rd("src/flow/flow_models/four-eqn_conservative/FlowModelFourEqnConservative.cpp")
    bu = rd("src/flow/flow_models/four-eqn_conservative/FlowModelBasicUtilitiesFourEqnConservative.cpp")
    mr = rd("src/util/mixing_rules/equations_of_state/EquationOfStateMixingRules.cpp")
    mi_all = rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateMixingRulesIdealGas.cpp")
    ig = rd("src/util/mixing_rules/equations_of_state/ideal_gas/EquationOfStateIdealGas.cpp")
    hpp = rd("include/flow/flow_models/four-eqn_conservative/FlowModelBasicUtilitiesFourEqnConservative.hpp")
    s_rho = statement(mr, r"rho\[idx_mixture_density\] \+= Z_rho\[si\]\[idx_partial_densities\]")
    s_Y = statement(fm, r"Y\[si\]\[idx_mass_fractions\] = rho_Y\[si\]\[idx\]/rho\[idx_density\]")
    s_u = statement(fm, r"u\[idx_velocity\] = rho_u\[idx\]/rho\[idx_density\]")
    s_v = statement(fm, r"v\[idx_velocity\] = rho_v\[idx\]/rho\[idx_density\]")
    s_w = statement(fm, r"w\[idx_velocity\] = rho_w\[idx\]/rho\[idx_density\]")
    s_e = statement(fm, r"epsilon\[idx_internal_energy\] = E\[idx\]/rho\[idx_density\] -\s*double\(1\)/double\(2\)\*\(u\[idx_velocity\]\*u\[idx_velocity\] \+ v\[idx_velocity\]\*v\[idx_velocity\] \+")
    ctor = line_range(mi_all, 100, 125)
    s_scp = statement(ctor, r"d_species_c_p\.push_back\(d_species_gamma\[si\]/\(d_species_gamma\[si\] - double\(1\)\)\*d_species_R\[si\]\)")
    s_scv = statement(ctor, r"d_species_c_v\.push_back\(double\(1\)/\(d_species_gamma\[si\] - double\(1\)\)\*d_species_R\[si\]\)")
    mi = line_range(mi_all, 6946, 7150)      # computeMixtureThermodynamicPropertiesWithMassFractions, all species given
    s_cp = statement(mi, r"c_p\[idx_mixture_thermo_properties\] \+= Y\[si\]\[idx_mass_fractions\]\*d_species_c_p\[si\]")
    s_cv = statement(mi, r"c_v\[idx_mixture_thermo_properties\] \+= Y\[si\]\[idx_mass_fractions\]\*d_species_c_v\[si\]")
    s_gm = statement(mi, r"gamma\[idx_mixture_thermo_properties\] = c_p\[idx_mixture_thermo_properties\]/")
    s_p = statement(ig, r"p\[idx_pressure\] = \(gamma\[idx_thermo_properties\] - double\(1\)\)\*rho\[idx_density\]\*\s*epsilon\[idx_internal_energy\]")
    s_ep = statement(ig, r"epsilon\[idx_internal_energy\] = p\[idx_pressure\]/\(\(gamma\[idx_thermo_properties\] - double\(1\)\)\*\s*rho\[idx_density\]\)")
    s_Gr = statement(ig, r"Gamma\[idx_gruneisen_parameter\] = gamma\[idx_thermo_properties\] - double\(1\)")
    ps = line_range(mi_all, 6367, 6564)
    s_Psi = statement(ps, r"Psi_i\[idx_partial_pressure_partial_partial_densities\] =")
    s_c0 = statement(fm, r"c\[idx_sound_speed\] = Gamma\[idx_sound_speed\]\*p\[idx_pressure\]/rho\[idx_density\]")
    s_c1 = statement(fm, r"c\[idx_sound_speed\] \+= Y\[si\]\[idx_mass_fractions\]\*Psi\[si\]\[idx_sound_speed\]")
    s_c2 = statement(fm, r"c\[idx_sound_speed\] = sqrt\(c\[idx_sound_speed\]\)")
    # bounds flag, 3-D x block of checkSideDataOfPrimitiveVariablesBounded
    bb = line_range(bu, 4380, 4510)
    s_brho = statement(bb, r"rho\[idx_face\] \+= V\[si\]\[idx_face\]", 0)
    s_bY = statement(bb, r"const double Y = V\[si\]\[idx_face\]/rho\[idx_face\]", 0)
    b_Y = if_else_block(bb, r"if \(Y > d_Y_bound_lo && Y < d_Y_bound_up\)", 0)
    b_r = if_else_block(bb, r"if \(rho\[idx_face\] > double\(0\)\)", 0)
    b_p = if_else_block(bb, r"if \(V\[d_num_species \+ d_dim\.getValue\(\)\]\[idx_face\] > double\(0\)\)", 0)

    def bound(nm):
        return re.search(nm + r" = double\(([-0-9.]+)\)", hpp).group(1)
    # projection / back-projection, 3-D x direction
    m = re.search(r"^FlowModelBasicUtilitiesFourEqnConservative::computeSideDataOfCharacteristicVariablesFromPrimitiveVariables\(", bu, flags=re.M)
    m2 = re.search(r"^FlowModelBasicUtilitiesFourEqnConservative::computeSideDataOfPrimitiveVariablesFromCharacteristicVariables\(", bu, flags=re.M)
    proj, back = bu[m.start():m2.start()], bu[m2.start():]
    s_za = statement(bu, r"rho_Y_average\[si\]\[idx_face_x\] = double\(1\)/double\(2\)\*\(rho_Y\[si\]\[idx_L\] \+ rho_Y\[si\]\[idx_R\]\)")
    s_ra = statement(bu, r"rho_average\[idx_face_x\] = double\(1\)/double\(2\)\*\(rho\[idx_density_L\] \+ rho\[idx_density_R\]\)")
    s_ca = statement(bu, r"c_average\[idx_face_x\] = double\(1\)/double\(2\)\*\(c\[idx_sound_speed_L\] \+ c\[idx_sound_speed_R\]\)")
    # the first statements that use V[d_num_species + 3][idx_p] are the 3-D x block's
    s_w0 = statement(proj, r"W\[0\]\[idx_face\] = V\[d_num_species\]\[idx_vel\] -\s*double\(1\)/\(rho_average\[idx_face\]\*c_average\[idx_face\]\)\*V\[d_num_species \+ 3\]\[idx_p\]", 0)
    s_wsi = statement(proj, r"W\[1 \+ si\]\[idx_face\] = V\[si\]\[idx_rho_Y\] - rho_Y_average\[si\]\[idx_face\]/\s*\(rho_average\[idx_face\]\*c_average\[idx_face\]\*c_average\[idx_face\]\)\*\s*V\[d_num_species \+ 3\]\[idx_p\]", 0)
    s_wl = statement(proj, r"W\[d_num_species \+ 3\]\[idx_face\] = V\[d_num_species\]\[idx_vel\] \+", 0)
    back3 = back[back.index("d_dim == tbox::Dimension(3)"):]
    s_vsi = statement(back3, r"V\[si\]\[idx_face\] = -double\(1\)/double\(2\)\*rho_Y_average", 0)
    s_vu = statement(back3, r"V\[d_num_species\]\[idx_face\] = double\(1\)/double\(2\)\*W\[0\]\[idx_face\] \+", 0)
    s_vp = statement(back3, r"V\[d_num_species \+ 3\]\[idx_face\] = -double\(1\)/double\(2\)\*rho_average", 0)
    return f"""
extern "C" void ref_path_points7(const double in[16], double out[15])
{{
    const int d_num_species = 2, d_num_eqn = 6;
    const std::vector<double> d_species_gamma = {{in[6], in[7]}}, d_species_R = {{in[8], in[9]}};
    std::vector<double> d_species_c_p, d_species_c_v;
    for (int si = 0; si < d_num_species; si++) {{ {s_scp} }}
    for (int si = 0; si < d_num_species; si++) {{ {s_scv} }}
    {{
        const int idx = 0, idx_density = 0, idx_mixture_density = 0, idx_partial_densities = 0, idx_mass_fractions = 0;
        const int idx_velocity = 0, idx_internal_energy = 0, idx_mixture_thermo_properties = 0;
        const int idx_thermo_properties = 0, idx_pressure = 0, idx_gruneisen_parameter = 0, idx_sound_speed = 0;
        const int idx_partial_pressure_partial_partial_densities = 0;
        double ry0[1] = {{in[0]}}, ry1[1] = {{in[1]}};
        double* rho_Y[2] = {{ry0, ry1}};
        double** Z_rho = rho_Y;            /* the mixing rules' generic name of the partial densities */
        const double rho_u[1] = {{in[2]}}, rho_v[1] = {{in[3]}}, rho_w[1] = {{in[4]}}, E[1] = {{in[5]}};
        double rho[1] = {{0.0}}, y0[1], y1[1], u[1], v[1], w[1], epsilon[1], gamma[1], p[1], Gamma[1], c[1];
        double c_p[1] = {{0.0}}, c_v[1] = {{0.0}}, psi0[1], psi1[1];
        double* Y[2] = {{y0, y1}};
        double* Psi[2] = {{psi0, psi1}};
        for (int si = 0; si < d_num_species; si++) {{ {s_rho} }}
        for (int si = 0; si < d_num_species; si++) {{ {s_Y} }}
        {s_u}
        {s_v}
        {s_w}
        {s_e}
        out[3] = epsilon[0];
        for (int si = 0; si < d_num_species; si++) {{
            {s_cp}
            {s_cv}
        }}
        {s_gm}
        {s_p}
        {s_ep}
        {s_Gr}
        for (int si = 0; si < d_num_species; si++) {{
            double* Psi_i = Psi[si];
            {s_Psi}
        }}
        {s_c0}
        for (int si = 0; si < d_num_species; si++) {{ {s_c1} }}
        {s_c2}
        out[0] = rho[0]; out[1] = y0[0]; out[2] = y1[0];
        out[4] = c_p[0]; out[5] = c_v[0]; out[6] = gamma[0]; out[7] = p[0]; out[8] = psi0[0]; out[9] = psi1[0]; out[10] = c[0];
    }}
    {{
        /* (rho, c, epsilon) of an interpolated side: the Riemann driver repeats the cell statements on side data */
        const int idx_density = 0, idx_mass_fractions = 0, idx_mixture_thermo_properties = 0, idx_thermo_properties = 0;
        const int idx_pressure = 0, idx_internal_energy = 0, idx_gruneisen_parameter = 0, idx_sound_speed = 0;
        const int idx_partial_pressure_partial_partial_densities = 0;
        const double* Vs = in + 10;
        double rho[1] = {{0.0}};
        for (int si = 0; si < d_num_species; si++) rho[0] += Vs[si];
        double y0[1] = {{Vs[0]/rho[0]}}, y1[1] = {{Vs[1]/rho[0]}};
        double* Y[2] = {{y0, y1}};
        double p[1] = {{Vs[d_num_species + 3]}}, c_p[1] = {{0.0}}, c_v[1] = {{0.0}}, gamma[1], epsilon[1], Gamma[1], c[1], psi0[1], psi1[1];
        double* Psi[2] = {{psi0, psi1}};
        for (int si = 0; si < d_num_species; si++) {{
            {s_cp}
            {s_cv}
        }}
        {s_gm}
        {s_ep}
        {s_Gr}
        for (int si = 0; si < d_num_species; si++) {{
            double* Psi_i = Psi[si];
            {s_Psi}
        }}
        {s_c0}
        for (int si = 0; si < d_num_species; si++) {{ {s_c1} }}
        {s_c2}
        out[11] = rho[0]; out[12] = c[0]; out[13] = epsilon[0];
    }}
    {{
        const int idx_face = 0;
        struct {{ int getValue() const {{ return 3; }} }} d_dim;
        const double d_Y_bound_lo = ({bound("d_Y_bound_lo")}), d_Y_bound_up = ({bound("d_Y_bound_up")});
        int are_bounded[1] = {{1}};
        double v_[6][1];
        double* V[6];
        for (int e = 0; e < 6; e++) {{ v_[e][0] = in[10 + e]; V[e] = v_[e]; }}
        double rho[1] = {{0.0}};
        for (int si = 0; si < d_num_species; si++) {{ {s_brho} }}
        for (int si = 0; si < d_num_species; si++) {{
            {s_bY}
            {b_Y}
        }}
        {b_r}
        {b_p}
        out[14] = (double)are_bounded[0];
    }}
}}

extern "C" void ref_path_points8(const double in[24], double out[16])
{{
    /* four-eqn conservative, 3-D x: face averages (in[0..7]: rhoY0 L/R, rhoY1 L/R, rho L/R, c L/R), projection of V
     * (in[8..13]) and back-projection of W (in[14..19]) with those averages */
    const int d_num_species = 2, d_num_eqn = 6;
    double rya0[1], rya1[1], rho_average[1], c_average[1];
    double* rho_Y_average[2] = {{rya0, rya1}};
    {{
        const int idx_face_x = 0, idx_L = 0, idx_R = 1, idx_density_L = 0, idx_density_R = 1;
        const int idx_sound_speed_L = 0, idx_sound_speed_R = 1;
        const double ry0[2] = {{in[0], in[1]}}, ry1[2] = {{in[2], in[3]}}, rho[2] = {{in[4], in[5]}}, c[2] = {{in[6], in[7]}};
        const double* rho_Y[2] = {{ry0, ry1}};
        for (int si = 0; si < d_num_species; si++) {{ {s_za} }}
        {s_ra}
        {s_ca}
        out[0] = rya0[0]; out[1] = rya1[0]; out[2] = rho_average[0]; out[3] = c_average[0];
    }}
    {{
        const int idx_face = 0, idx_rho_Y = 0, idx_vel = 0, idx_p = 0;
        double v_[6][1], w_[6][1];
        double *V[6], *W[6];
        for (int e = 0; e < 6; e++) {{ v_[e][0] = in[8 + e]; V[e] = v_[e]; W[e] = w_[e]; }}
        for (int si = 0; si < d_num_species; si++) {{ {s_wsi} }}
        {s_w0}
        W[d_num_species + 1][idx_face] = V[d_num_species + 1][idx_vel];
        W[d_num_species + 2][idx_face] = V[d_num_species + 2][idx_vel];
        {s_wl}
        for (int e = 0; e < 6; e++) out[4 + e] = W[e][0];
    }}
    {{
        const int idx_face = 0;
        double v_[6][1], w_[6][1];
        double *V[6], *W[6];
        for (int e = 0; e < 6; e++) {{ w_[e][0] = in[14 + e]; V[e] = v_[e]; W[e] = w_[e]; }}
        for (int si = 0; si < d_num_species; si++) {{ {s_vsi} }}
        {s_vu}
        V[d_num_species + 1][idx_face] = W[d_num_species + 1][idx_face];
        V[d_num_species + 2][idx_face] = W[d_num_species + 2][idx_face];
        {s_vp}
        for (int e = 0; e < 6; e++) out[10 + e] = V[e][0];
    }}
}}
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
    parts.append(path_statements2())
    parts.append(path_statements3())
    parts.append(path_statements4())
    parts.append(path_statements5())
    parts.append(path_statements6())
    parts.append(FC_WRAPPER)
    parts.append(path_statements7())
    parts.append(diffusive_kernels())
    parts.append(diffusive_term_tables())
    parts.append(diffusive_dt_statements())
    parts.append(midpoint_kernels())
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
