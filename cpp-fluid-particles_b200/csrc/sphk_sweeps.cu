// sphk_sweeps.cu -- the per-particle neighbour sweeps of the three solvers + element-wise steps.
//
// One generic sweep driver, instantiated per operator:
//   * k_sweep_cells<Op> : walks the 27 neighbour cells in the reference's order (x outermost, z
//     innermost; inside a cell first the fluid range, then the boundary range -- SURVEY 2.2), reading
//     packed float4 {x,y,z,mass} records (one LDG.128 per candidate instead of the reference's four
//     scalar loads) and accumulating in REGISTERS (the reference accumulates density straight into
//     global memory, BasicSPHSolver.cu:37,48).
//   * k_sweep_list<Op>  : walks a per-step neighbour list (built once per neighbour search by the
//     same cell walk, candidates kept in the reference's order) while positions are unchanged -- the
//     ~85% of candidate pairs outside the support are tested once per step instead of once per sweep
//     (DFSPH runs 23 sweeps per step on identical positions).
// Because the candidate ORDER is the reference's in both drivers and the operators evaluate the same
// expressions with the same fast-math intrinsics (this file is compiled with -use_fast_math like the
// reference, Q10), results agree with the reference kernels to a few ulp, far inside 1e-5.
#include <cstdio>
#include "sphk_internal.cuh"

// =================================================================================================
// Operators.  Acc = per-particle register accumulator; pair<B>() adds neighbour j (B: boundary).
// j is the unified index (boundary particle b is capF + b); pj.w is mass_j.
// =================================================================================================

// computeDensity_CUDA, BasicSPHSolver.cu:32-72
struct OpDensity {
    float* density;
    static constexpr bool kFluidOnly = false, kSplitB = false;
    struct Acc { float rho; };
    struct Nb {};
    __device__ Nb fetch(int, bool) const { return Nb{}; }
    __device__ void begin(Acc& a, int, float4, const DevScene&) const { a.rho = 0.f; }
    template <bool B> __device__ void pair(Acc& a, int, int, float3, float r2, float4 pj, Nb, const DevScene& s) const {
        a.rho += pj.w * w_cubic(sqrtf(r2), s.R);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const { density[i] = a.rho; }
};

// computeBoundaryMass_CUDA, SPHSystem.cu:79-105 (driver: k_boundary_mass below)

// pressureForce_CUDA, BasicSPHSolver.cu:113-165.  prho[j] = p_j / max(eps, rho_j^2) precomputed per
// particle (identical value to the reference's per-pair expression); boundary part is zero.
struct OpPressureForce {
    const float* prho; float4* vel4; float* vel; float dt;
    static constexpr bool kFluidOnly = false, kSplitB = true;
    struct Acc { float3 a; float pri; };
    struct Nb { float prj; };
    __device__ Nb fetch(int j, bool isB) const { return Nb{isB ? 0.f : prho[j]}; }
    __device__ void begin(Acc& a, int i, float4, const DevScene&) const { a.a = f3(0, 0, 0); a.pri = prho[i]; }
    template <bool B> __device__ void pair(Acc& a, int i, int j, float3 d, float r2, float4 pj, Nb nb, const DevScene& s) const {
        if (B) a.a += -pj.w * a.pri * grad_w_cubic(d, sqrtf(r2), s.R);
        else if (i != j) a.a += -pj.w * (a.pri + nb.prj) * grad_w_cubic(d, sqrtf(r2), s.R);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        float3 acc = a.a;
        const float l2 = dot3(acc, acc);
        if (sqrtf(l2) > SPHK_MAX_A) acc = acc * rsqrtf(l2) * SPHK_MAX_A;      // :160-161
        float4 v = vel4[i];
        v.x += acc.x * dt; v.y += acc.y * dt; v.z += acc.z * dt;
        vel4[i] = v; store3(vel, i, xyz(v));
    }
};

// viscosity_CUDA + vel += deltaV, BasicSPHSolver.cu:183-225
struct OpViscosity {
    const float4* vel4_in; float4* vel4_out; float* vel; float* deltaV; float rho0, visc, dt;
    static constexpr bool kFluidOnly = true, kSplitB = false;
    struct Acc { float3 a; float3 vi; };
    struct Nb { float4 v; };
    __device__ Nb fetch(int j, bool isB) const { return Nb{isB ? make_float4(0, 0, 0, 0) : vel4_in[j]}; }
    __device__ void begin(Acc& a, int i, float4, const DevScene&) const { a.a = f3(0, 0, 0); a.vi = xyz(vel4_in[i]); }
    template <bool B> __device__ void pair(Acc& a, int, int, float3, float r2, float4 pj, Nb nb, const DevScene& s) const {
        const float3 vj = xyz(nb.v);
        a.a += pj.w * ((vj - a.vi) / rho0) * lap_visc(sqrtf(r2), s.R);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        const float3 dv = visc * a.a * dt;
        store3(deltaV, i, dv);
        const float3 v = a.vi + dv;
        vel4_out[i] = make_float4(v.x, v.y, v.z, 0.f); store3(vel, i, v);
    }
};

// computeColorGrad_CUDA, BasicSPHSolver.cu:277-318
struct OpColorGrad {
    float* colorGrad; float rho0, rhoB;
    static constexpr bool kFluidOnly = false, kSplitB = true;
    struct Acc { float3 num; float den; };
    struct Nb {};
    __device__ Nb fetch(int, bool) const { return Nb{}; }
    __device__ void begin(Acc& a, int, float4, const DevScene&) const { a.num = f3(0, 0, 0); a.den = 0.f; }
    template <bool B> __device__ void pair(Acc& a, int, int, float3 d, float r2, float4 pj, Nb, const DevScene& s) const {
        const float r = sqrtf(r2);
        const float V = pj.w / (B ? rhoB : rho0);
        a.num += V * grad_w_cubic(d, r, s.R);
        a.den += V * w_cubic(r, s.R);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const { store3(colorGrad, i, a.num / fmaxf(SPHK_EPS, a.den)); }
};

// surfaceTensionAndAirPressure_CUDA, BasicSPHSolver.cu:332-370.  cg2[j] = dot(c_j, c_j) precomputed.
struct OpSurface {
    const float* colorGrad; const float* cg2; float4* vel4; float* vel; float dt, rho0, kappa, airP;
    static constexpr bool kFluidOnly = true, kSplitB = false;
    struct Acc { float3 a; float cii, lci; };
    __device__ void begin(Acc& a, int i, float4, const DevScene&) const {
        a.a = f3(0, 0, 0);
        const float3 ci = load3(colorGrad, i);
        a.cii = dot3(ci, ci); a.lci = sqrtf(a.cii);
    }
    struct Nb { float cjj; };
    __device__ Nb fetch(int j, bool isB) const { return Nb{isB ? 0.f : cg2[j]}; }
    template <bool B> __device__ void pair(Acc& a, int, int, float3 d, float r2, float4 pj, Nb nb, const DevScene& s) const {
        const float r = sqrtf(r2);
        a.a += 0.25f * pj.w / (rho0 * rho0) * kappa * (a.cii + nb.cjj) * grad_surface_tension(d, r, s.R);
        a.a += airP * pj.w / (rho0 * rho0) * grad_w_cubic(d, r, s.R) * a.lci / fmaxf(SPHK_EPS, a.lci);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        float4 v = vel4[i];
        v.x += a.a.x * dt; v.y += a.a.y * dt; v.z += a.a.z * dt;
        vel4[i] = v; store3(vel, i, xyz(v));
    }
};

// computeDensityAlpha_CUDA, DFSPHSolver.cu:212-249
struct OpDensityAlpha {
    float* density; float* alpha;
    static constexpr bool kFluidOnly = false, kSplitB = true;
    struct Acc { float den, lam; float3 gs; };
    struct Nb {};
    __device__ Nb fetch(int, bool) const { return Nb{}; }
    __device__ void begin(Acc& a, int, float4, const DevScene&) const { a.den = 0.f; a.lam = 0.f; a.gs = f3(0, 0, 0); }
    template <bool B> __device__ void pair(Acc& a, int, int, float3 d, float r2, float4 pj, Nb, const DevScene& s) const {
        const float r = sqrtf(r2);
        a.den += pj.w * w_cubic(r, s.R);
        const float3 mg = pj.w * grad_w_cubic(d, r, s.R);
        a.gs += mg;
        if (!B) a.lam += dot3(mg, mg);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        density[i] = a.den;
        alpha[i] = -1.0f / fmaxf(SPHK_EPS, dot3(a.gs, a.gs) + a.lam);
    }
};

// computeDivergenceError_CUDA (DFSPHSolver.cu:261-306, kDensity=false) and computeDensityError_CUDA
// (:74-116, kDensity=true; optionally with the warm-stiffness accumulate of :199-203 fused)
template <bool kDensity> struct OpDfsphError {
    const float4* vel4; const float* density; const float* alpha; float* error; float* stiff; float* warm;
    float dt, rho0;
    static constexpr bool kFluidOnly = false, kSplitB = true;
    struct Acc { float e; float3 vi; };
    __device__ void begin(Acc& a, int i, float4, const DevScene&) const { a.e = 0.f; a.vi = xyz(vel4[i]); }
    struct Nb { float4 v; };
    __device__ Nb fetch(int j, bool isB) const { return Nb{isB ? make_float4(0, 0, 0, 0) : vel4[j]}; }
    template <bool B> __device__ void pair(Acc& a, int, int, float3 d, float r2, float4 pj, Nb nb, const DevScene& s) const {
        const float3 g = grad_w_cubic(d, sqrtf(r2), s.R);
        if (B) a.e += pj.w * dot3(a.vi, g);
        else a.e += pj.w * dot3(a.vi - xyz(nb.v), g);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        float e;
        if (kDensity) e = fmaxf(0.0f, dt * a.e + density[i] - rho0);
        else {
            e = fmaxf(0.0f, a.e);
            if (density[i] + dt * e < rho0 && density[i] <= rho0) e = 0.0f;     // :302-303
        }
        error[i] = e;
        const float k = e * alpha[i];
        stiff[i] = k;
        if (kDensity && warm) warm[i] += k;
    }
};

// correctDivergenceError_CUDA (DFSPHSolver.cu:308-329) / correctDensityError_CUDA (:118-158, kDivDt)
// also computeDeltaPos_CUDA (PBDSolver.cu:170-210) shares the pair term with lambda as the scalar.
template <int kMode /*0: vel += a, 1: vel += a/dt, 2: deltaPos = a/rho0*/> struct OpScalarGradient {
    const float* kappa; float4* vel4; float* vel; float* deltaPos; float dt_or_rho0;
    static constexpr bool kFluidOnly = false, kSplitB = true;
    struct Acc { float3 a; float ki; };
    __device__ void begin(Acc& a, int i, float4, const DevScene&) const { a.a = f3(0, 0, 0); a.ki = kappa[i]; }
    struct Nb { float kj; };
    __device__ Nb fetch(int j, bool isB) const { return Nb{isB ? 0.f : kappa[j]}; }
    template <bool B> __device__ void pair(Acc& a, int, int, float3 d, float r2, float4 pj, Nb nb, const DevScene& s) const {
        const float3 g = grad_w_cubic(d, sqrtf(r2), s.R);
        if (B) a.a += pj.w * a.ki * g;
        else a.a += pj.w * (a.ki + nb.kj) * g;
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        if (kMode == 2) { store3(deltaPos, i, a.a / dt_or_rho0); return; }
        const float3 dv = (kMode == 1) ? a.a / dt_or_rho0 : a.a;
        float4 v = vel4[i];
        v.x += dv.x; v.y += dv.y; v.z += dv.z;
        vel4[i] = v; store3(vel, i, xyz(v));
    }
};

// computeDensityLambda_CUDA, PBDSolver.cu:127-168 (rho0 passed through `bool`, Q4)
struct OpPbdLambda {
    float* density; float* lambda; float rho0, rho0AsBool, relaxation;
    static constexpr bool kFluidOnly = false, kSplitB = false;
    struct Acc { float den, lam; float3 gs; };
    struct Nb {};
    __device__ Nb fetch(int, bool) const { return Nb{}; }
    __device__ void begin(Acc& a, int, float4, const DevScene&) const { a.den = 0.f; a.lam = 0.f; a.gs = f3(0, 0, 0); }
    template <bool B> __device__ void pair(Acc& a, int, int, float3 d, float r2, float4 pj, Nb, const DevScene& s) const {
        const float r = sqrtf(r2);
        a.den += pj.w * w_cubic(r, s.R);
        const float3 g = -pj.w * grad_w_cubic(d, r, s.R) / rho0AsBool;
        a.gs -= g;
        a.lam += dot3(g, g);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        density[i] = a.den;
        float l = (a.den > rho0) ? (-(a.den / rho0 - 1.0f) / (dot3(a.gs, a.gs) + a.lam + SPHK_EPS)) : 0.0f;
        lambda[i] = l * relaxation;
    }
};

// XSPHViscosity_CUDA, PBDSolver.cu:89-115, Jacobi: reads vel4_in, writes vel4_out (Q5)
struct OpXsph {
    const float4* vel4_in; float4* vel4_out; float c, rho0;
    static constexpr bool kFluidOnly = true, kSplitB = false;
    struct Acc { float3 a; float3 vi; };
    __device__ void begin(Acc& a, int i, float4, const DevScene&) const { a.a = f3(0, 0, 0); a.vi = xyz(vel4_in[i]); }
    struct Nb { float4 v; };
    __device__ Nb fetch(int j, bool isB) const { return Nb{isB ? make_float4(0, 0, 0, 0) : vel4_in[j]}; }
    template <bool B> __device__ void pair(Acc& a, int, int, float3, float r2, float4 pj, Nb nb, const DevScene& s) const {
        a.a += pj.w * (xyz(nb.v) - a.vi) * w_cubic(sqrtf(r2), s.R);
    }
    __device__ void end(Acc& a, int i, float4, const DevScene&) const {
        const float3 v = a.vi + c * a.a / rho0;
        vel4_out[i] = make_float4(v.x, v.y, v.z, 0.f);
    }
};

// neighbour-list builder: keeps every candidate with r^2 <= r2cut in the cell walk's order, self excluded
// (the self pair contributes exactly 0 to every operator: W(0)=0 by Q1, grad W(0)=0, v_i - v_i = 0).
struct OpBuildList {
    int* nbr; int* cnt;
    static constexpr bool kFluidOnly = false, kSplitB = false;
    struct Acc { int n; };
    struct Nb {};
    __device__ Nb fetch(int, bool) const { return Nb{}; }
    __device__ void begin(Acc& a, int, float4, const DevScene&) const { a.n = 0; }
    // layout: entries 4b..4b+3 of particle i form the int4 at nbr4[b * stride + i] (one coalesced LDG.128 per
    // batch of four neighbours in the walk)
    __device__ size_t slot(int n, int i, const DevScene& s) const {
        return (static_cast<size_t>(n >> 2) * s.nbrStride + i) * 4 + (n & 3);
    }
    template <bool B> __device__ void pair(Acc& a, int i, int j, float3, float, float4, Nb, const DevScene& s) const {
        if (j == i) return;
        if (a.n < s.kmax) nbr[slot(a.n, i, s)] = j;
        ++a.n;
    }
    __device__ void end(Acc& a, int i, float4, const DevScene& s) const {
        cnt[i] = a.n;
        // pad the last batch with the particle itself: the self pair contributes exactly 0 to every operator
        for (int n = a.n; (n & 3) && n < s.kmax; ++n) nbr[slot(n, i, s)] = i;
    }
};

// =================================================================================================
// Sweep drivers
// =================================================================================================
template <class Op>
__device__ __forceinline__ void walk_cells(const DevScene& s, const Op& op, typename Op::Acc& acc, int i, float4 pi) {
    const int cx = cell_coord(pi.x, s.cellLength), cy = cell_coord(pi.y, s.cellLength), cz = cell_coord(pi.z, s.cellLength);
    const float3 xi = xyz(pi);
#pragma unroll 1
    for (int m = 0; m < 27; ++m) {
        const int c = cell_index(cx + m / 9 - 1, cy + (m % 9) / 3 - 1, cz + m % 3 - 1, s.cs);
        if (c == s.cs.x * s.cs.y * s.cs.z) continue;
        {
            const int end = s.csF[c + 1];
            for (int j = s.csF[c]; j < end; ++j) {
                const float4 pj = s.posm[j];
                const float3 d = xi - xyz(pj);
                const float r2 = dot3(d, d);
                if (r2 <= s.r2cut) op.template pair<false>(acc, i, j, d, r2, pj, op.fetch(j, false), s);
            }
        }
        if (!Op::kFluidOnly) {
            const int end = s.csB[c + 1];
            for (int jb = s.csB[c]; jb < end; ++jb) {
                const int j = s.bOff + jb;
                const float4 pj = s.posm[j];
                const float3 d = xi - xyz(pj);
                const float r2 = dot3(d, d);
                if (r2 <= s.r2cut) op.template pair<true>(acc, i, j, d, r2, pj, op.fetch(j, true), s);
            }
        }
    }
}

template <class Op>
__global__ void __launch_bounds__(SPHK_BLOCK) k_sweep_cells(const DevScene s, const Op op) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= s.nF) return;
    const float4 pi = s.posm[i];
    typename Op::Acc acc;
    op.begin(acc, i, pi, s);
    walk_cells(s, op, acc, i, pi);
    op.end(acc, i, pi, s);
}

template <class Op>
__device__ __forceinline__ void list_pair(const DevScene& s, const Op& op, typename Op::Acc& acc, int i, float3 xi, int j,
                                          float4 pj, typename Op::Nb nb) {
    const bool isB = j >= s.bOff;
    if (Op::kFluidOnly && isB) return;
    const float3 d = xi - xyz(pj);
    const float r2 = dot3(d, d);
    if (Op::kSplitB && isB) op.template pair<true>(acc, i, j, d, r2, pj, nb, s);
    else op.template pair<false>(acc, i, j, d, r2, pj, nb, s);
}

// List walk, software-pipelined for memory-level parallelism: the four indices of a batch arrive in ONE
// coalesced 16-byte load (streaming, evict-first: the list is read once per sweep and must not push the
// gathered particle records out of L2), the next batch's indices are prefetched before the current batch
// is consumed, and the 4 (+4 payload) dependent gathers of a batch are issued back to back before any math.
template <class Op>
__global__ void __launch_bounds__(SPHK_BLOCK) k_sweep_list(const DevScene s, const Op op) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= s.nF) return;
    const float4 pi = s.posm[i];
    const float3 xi = xyz(pi);
    typename Op::Acc acc;
    op.begin(acc, i, pi, s);
    const int n = s.cnt[i];
    if (n <= s.kmax) {
        const int nb4 = (n + 3) >> 2;
        const int4* __restrict__ row = reinterpret_cast<const int4*>(s.nbr) + i;
        int4 jn = make_int4(i, i, i, i);
        if (nb4 > 0) jn = __ldcs(row);
        for (int b = 0; b < nb4; ++b) {
            const int4 j4 = jn;
            row += s.nbrStride;
            if (b + 1 < nb4) jn = __ldcs(row);
            const float4 p0 = __ldg(s.posm + j4.x), p1 = __ldg(s.posm + j4.y), p2 = __ldg(s.posm + j4.z), p3 = __ldg(s.posm + j4.w);
            const typename Op::Nb n0 = op.fetch(j4.x, j4.x >= s.bOff), n1 = op.fetch(j4.y, j4.y >= s.bOff),
                                  n2 = op.fetch(j4.z, j4.z >= s.bOff), n3 = op.fetch(j4.w, j4.w >= s.bOff);
            list_pair(s, op, acc, i, xi, j4.x, p0, n0);
            list_pair(s, op, acc, i, xi, j4.y, p1, n1);
            list_pair(s, op, acc, i, xi, j4.z, p2, n2);
            list_pair(s, op, acc, i, xi, j4.w, p3, n3);
        }
    } else {
        walk_cells(s, op, acc, i, pi);      // more neighbours than the list keeps: exact fallback
    }
    op.end(acc, i, pi, s);
}

// computeBoundaryMass_CUDA, SPHSystem.cu:79-105: boundary particles against the boundary set only
__global__ void __launch_bounds__(SPHK_BLOCK)
k_boundary_mass(float4* __restrict__ posmB, float* __restrict__ mass, int n, const int* __restrict__ csB, int3 cs,
                float cellLength, float rhoB, float R) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    const float4 pi = posmB[i];
    const int cx = cell_coord(pi.x, cellLength), cy = cell_coord(pi.y, cellLength), cz = cell_coord(pi.z, cellLength);
    float sum = 0.f;
    for (int m = 0; m < 27; ++m) {
        const int c = cell_index(cx + m / 9 - 1, cy + (m % 9) / 3 - 1, cz + m % 3 - 1, cs);
        if (c == cs.x * cs.y * cs.z) continue;
        const int end = csB[c + 1];
        for (int j = csB[c]; j < end; ++j) {
            const float3 d = xyz(pi) - xyz(posmB[j]);
            sum += w_cubic(sqrtf(dot3(d, d)), R);
        }
    }
    mass[i] = rhoB / fmaxf(SPHK_EPS, sum);
}
__global__ void __launch_bounds__(SPHK_BLOCK) k_set_w(float4* __restrict__ posm, const float* __restrict__ mass, int n) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i < n) posm[i].w = mass[i];
}

// ---- element-wise kernels -------------------------------------------------------------------------
__global__ void __launch_bounds__(SPHK_BLOCK)
k_gravity(float4* __restrict__ vel4, float* __restrict__ vel, int n, float3 dv) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    float4 v = vel4[i];
    v.x += dv.x; v.y += dv.y; v.z += dv.z;
    vel4[i] = v; store3(vel, i, xyz(v));
}

// computePressure_CUDA, BasicSPHSolver.cu:103-111
__global__ void __launch_bounds__(SPHK_BLOCK)
k_pressure(const float* __restrict__ density, float* __restrict__ pressure, int n, float rho0, float stiff) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    float p = stiff * (powf((density[i] / rho0), 7) - 1.0f);
    if (p < 0.0f) p = 0.0f;
    pressure[i] = p;
}
__global__ void __launch_bounds__(SPHK_BLOCK)
k_prho(const float* __restrict__ density, const float* __restrict__ pressure, float* __restrict__ prho, int n) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i < n) prho[i] = pressure[i] / fmaxf(SPHK_EPS, density[i] * density[i]);
}
__global__ void __launch_bounds__(SPHK_BLOCK) k_cg2(const float* __restrict__ cg, float* __restrict__ cg2, int n) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    const float3 c = load3(cg, i);
    cg2[i] = dot3(c, c);
}

// Particles::advect (Particles.cu:28-36) + enforceBoundary_CUDA(pos, vel) (BasicSPHSolver.cu:85-96)
__device__ __forceinline__ void clamp_axis(float& p, float* v, float L) {
    if (p <= L * .00f) { p = L * .00f; if (v) *v = fmaxf(*v, 0.0f); }
    if (p >= L * .99f) { p = L * .99f; if (v) *v = fminf(*v, 0.0f); }
}
__global__ void __launch_bounds__(SPHK_BLOCK)
k_advect(float4* __restrict__ posm, float4* __restrict__ vel4, float* __restrict__ pos, float* __restrict__ vel, int n,
         float dt, float3 space) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    float4 p = posm[i];
    float4 v = vel4[i];
    p.x = p.x + dt * v.x; p.y = p.y + dt * v.y; p.z = p.z + dt * v.z;
    clamp_axis(p.x, &v.x, space.x); clamp_axis(p.y, &v.y, space.y); clamp_axis(p.z, &v.z, space.z);
    posm[i] = p; vel4[i] = v;
    store3(pos, i, xyz(p)); store3(vel, i, xyz(v));
}
// thrust::transform(pos += dpos) + enforceBoundary_CUDA(pos), PBDSolver.cu:212-223,247-253
__global__ void __launch_bounds__(SPHK_BLOCK)
k_apply_delta_pos(float4* __restrict__ posm, float* __restrict__ pos, const float* __restrict__ dpos, int n, float3 space) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    float4 p = posm[i];
    const float3 d = load3(dpos, i);
    p.x += d.x; p.y += d.y; p.z += d.z;
    clamp_axis(p.x, nullptr, space.x); clamp_axis(p.y, nullptr, space.y); clamp_axis(p.z, nullptr, space.z);
    posm[i] = p; store3(pos, i, xyz(p));
}
// vel = (pos - posLast) / dt, PBDSolver.cu:55-60
__global__ void __launch_bounds__(SPHK_BLOCK)
k_vel_from_pos(const float4* __restrict__ posm, const float* __restrict__ posLast, float4* __restrict__ vel4,
               float* __restrict__ vel, int n, float dt) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    const float3 v = (xyz(posm[i]) - load3(posLast, i)) / dt;
    vel4[i] = make_float4(v.x, v.y, v.z, 0.f); store3(vel, i, v);
}
__global__ void __launch_bounds__(SPHK_BLOCK)
k_commit_vel(const float4* __restrict__ src, float4* __restrict__ vel4, float* __restrict__ vel, int n) {
    const int i = blockIdx.x * SPHK_BLOCK + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    vel4[i] = v; store3(vel, i, xyz(v));
}
__global__ void __launch_bounds__(256) k_list_stats(const int* __restrict__ cnt, int n, int kmax, unsigned long long* out) {
    unsigned long long mx = 0, ov = 0, tot = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int c = cnt[i];
        if (c > mx) mx = c;
        ov += c > kmax; tot += c;
    }
    atomicMax(out, mx); atomicAdd(out + 1, ov); atomicAdd(out + 2, tot);
}

// =================================================================================================
// Host side
// =================================================================================================
static int check_scene(const sphk_ctx* c, const sphk_scene* s) {
    if (!c || !s) return SPHK_ERR_INVALID;
    if (!c->fluidSearched || !c->boundarySearched) return SPHK_ERR_STATE;
    if (s->fluid.n != c->nF || s->boundary.n != c->nB) return SPHK_ERR_STATE;
    if (!s->cell_start_fluid || !s->cell_start_boundary || !s->fluid.pos || !s->fluid.vel) return SPHK_ERR_INVALID;
    if (!(s->radius > 0.f)) return SPHK_ERR_INVALID;
    return SPHK_OK;
}

static DevScene dev_scene(const sphk_ctx* c, const sphk_scene* s) {
    DevScene d;
    d.posm = c->posm; d.csF = s->cell_start_fluid; d.csB = s->cell_start_boundary;
    d.nbr = c->nbr; d.cnt = c->cnt;
    d.nF = c->nF; d.bOff = c->capF; d.nbrStride = c->capF; d.kmax = c->kmax;
    d.cs = c->cs; d.cellLength = c->cellLength; d.R = s->radius;
    // candidates beyond the support contribute exactly 0 to every operator; the margin only covers the
    // approximate sqrt/div of the fast-math support tests (q > 2, r <= R)
    d.r2cut = s->radius * s->radius * (1.0f + 1e-5f);
    return d;
}

static int ensure_list(sphk_ctx* c, const DevScene& d) {
    if (c->listEpoch == c->searchEpoch) return SPHK_OK;
    if (!c->nbr) {
        const size_t bytes = sizeof(int) * static_cast<size_t>(c->kmax) * static_cast<size_t>(c->capF);
        if (cudaMalloc(reinterpret_cast<void**>(&c->nbr), bytes) != cudaSuccess) { cudaGetLastError(); return SPHK_ERR_ALLOC; }
    }
    DevScene b = d;
    b.nbr = c->nbr;
    OpBuildList op{c->nbr, c->cnt};
    k_sweep_cells<OpBuildList><<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(b, op);
    c->launches++;
    c->listEpoch = c->searchEpoch;
    return SPHK_OK;
}

template <class Op> static int run_sweep(sphk_ctx* c, const sphk_scene* s, const Op& op) {
    DevScene d = dev_scene(c, s);
    const bool list = c->useList && !c->posDirty;
    if (list) {
        const int rc = ensure_list(c, d);
        if (rc != SPHK_OK) return rc;
        d.nbr = c->nbr;
        k_sweep_list<Op><<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(d, op);
    } else {
        k_sweep_cells<Op><<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(d, op);
    }
    c->launches++;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

#define SPHK_CHECK_SCENE(c, s) do { const int rc_ = check_scene((c), (s)); if (rc_ != SPHK_OK) return rc_; } while (0)

extern "C" int sphk_boundary_mass(sphk_ctx* c, const sphk_particles* b, const int* csB, float rhoB, float R) {
    if (!c || !b || !csB || !b->mass) return SPHK_ERR_INVALID;
    if (!c->boundarySearched || b->n != c->nB) return SPHK_ERR_STATE;
    float4* posmB = c->posm + c->capF;
    k_boundary_mass<<<sphk_blocks(c->nB), SPHK_BLOCK, 0, c->stream>>>(posmB, b->mass, c->nB, csB, c->cs, c->cellLength, rhoB, R);
    k_set_w<<<sphk_blocks(c->nB), SPHK_BLOCK, 0, c->stream>>>(posmB, b->mass, c->nB);
    c->launches += 2;
    c->listEpoch = ~0ull;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_gravity(sphk_ctx* c, const sphk_scene* s, float dt, const float G[3]) {
    SPHK_CHECK_SCENE(c, s);
    if (!G) return SPHK_ERR_INVALID;
    const float3 dv = make_float3(dt * G[0], dt * G[1], dt * G[2]);   // const auto dv = dt * G, BasicSPHSolver.cu:229
    k_gravity<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(c->vel4, s->fluid.vel, c->nF, dv);
    c->launches++;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_viscosity(sphk_ctx* c, const sphk_scene* s, float* delta_v, float rho0, float visc, float dt) {
    SPHK_CHECK_SCENE(c, s);
    if (!delta_v) return SPHK_ERR_INVALID;
    // Jacobi: every thread reads neighbours' OLD velocity (the reference writes deltaV to a buffer and adds
    // afterwards); results go to a temp and are committed by a second pass.
    float4* tmp = c->snapB;
    OpViscosity op{c->vel4, tmp, s->fluid.vel, delta_v, rho0, visc, dt};
    const int rc = run_sweep(c, s, op);
    if (rc != SPHK_OK) return rc;
    SPHK_CUDA_TRY(cudaMemcpyAsync(c->vel4, tmp, sizeof(float4) * static_cast<size_t>(c->nF), cudaMemcpyDeviceToDevice, c->stream));
    return SPHK_OK;
}

extern "C" int sphk_color_grad(sphk_ctx* c, const sphk_scene* s, float* color_grad, float rho0, float rhoB) {
    SPHK_CHECK_SCENE(c, s);
    if (!color_grad) return SPHK_ERR_INVALID;
    OpColorGrad op{color_grad, rho0, rhoB};
    return run_sweep(c, s, op);
}

extern "C" int sphk_surface(sphk_ctx* c, const sphk_scene* s, const float* color_grad, float dt, float rho0,
                            float kappa, float airP) {
    SPHK_CHECK_SCENE(c, s);
    if (!color_grad) return SPHK_ERR_INVALID;
    k_cg2<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(color_grad, c->aux, c->nF);
    c->launches++;
    OpSurface op{color_grad, c->aux, c->vel4, s->fluid.vel, dt, rho0, kappa, airP};
    return run_sweep(c, s, op);
}

extern "C" int sphk_density(sphk_ctx* c, const sphk_scene* s) {
    SPHK_CHECK_SCENE(c, s);
    if (!s->fluid.density) return SPHK_ERR_INVALID;
    OpDensity op{s->fluid.density};
    return run_sweep(c, s, op);
}

extern "C" int sphk_pressure(sphk_ctx* c, const sphk_scene* s, float rho0, float stiff) {
    SPHK_CHECK_SCENE(c, s);
    if (!s->fluid.density || !s->fluid.pressure) return SPHK_ERR_INVALID;
    k_pressure<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(s->fluid.density, s->fluid.pressure, c->nF, rho0, stiff);
    c->launches++;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_pressure_force(sphk_ctx* c, const sphk_scene* s, float dt) {
    SPHK_CHECK_SCENE(c, s);
    if (!s->fluid.density || !s->fluid.pressure) return SPHK_ERR_INVALID;
    k_prho<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(s->fluid.density, s->fluid.pressure, c->aux, c->nF);
    c->launches++;
    OpPressureForce op{c->aux, c->vel4, s->fluid.vel, dt};
    return run_sweep(c, s, op);
}

extern "C" int sphk_advect(sphk_ctx* c, const sphk_scene* s, float dt, const float space[3]) {
    SPHK_CHECK_SCENE(c, s);
    if (!space) return SPHK_ERR_INVALID;
    k_advect<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(c->posm, c->vel4, s->fluid.pos, s->fluid.vel, c->nF, dt,
                                                              make_float3(space[0], space[1], space[2]));
    c->launches++;
    c->posDirty = true;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_dfsph_density_alpha(sphk_ctx* c, const sphk_scene* s, float* alpha) {
    SPHK_CHECK_SCENE(c, s);
    if (!alpha || !s->fluid.density) return SPHK_ERR_INVALID;
    OpDensityAlpha op{s->fluid.density, alpha};
    return run_sweep(c, s, op);
}

extern "C" int sphk_dfsph_div_error(sphk_ctx* c, const sphk_scene* s, const float* alpha, float* error, float* stiff,
                                    float dt, float rho0) {
    SPHK_CHECK_SCENE(c, s);
    if (!alpha || !error || !stiff || !s->fluid.density) return SPHK_ERR_INVALID;
    OpDfsphError<false> op{c->vel4, s->fluid.density, alpha, error, stiff, nullptr, dt, rho0};
    return run_sweep(c, s, op);
}

extern "C" int sphk_dfsph_den_error(sphk_ctx* c, const sphk_scene* s, const float* alpha, float* error, float* stiff,
                                    float dt, float rho0, float* warm) {
    SPHK_CHECK_SCENE(c, s);
    if (!alpha || !error || !stiff || !s->fluid.density) return SPHK_ERR_INVALID;
    OpDfsphError<true> op{c->vel4, s->fluid.density, alpha, error, stiff, warm, dt, rho0};
    return run_sweep(c, s, op);
}

extern "C" int sphk_dfsph_div_correct(sphk_ctx* c, const sphk_scene* s, const float* stiff) {
    SPHK_CHECK_SCENE(c, s);
    if (!stiff) return SPHK_ERR_INVALID;
    OpScalarGradient<0> op{stiff, c->vel4, s->fluid.vel, nullptr, 1.0f};
    return run_sweep(c, s, op);
}

extern "C" int sphk_dfsph_den_correct(sphk_ctx* c, const sphk_scene* s, const float* stiff, float dt) {
    SPHK_CHECK_SCENE(c, s);
    if (!stiff) return SPHK_ERR_INVALID;
    OpScalarGradient<1> op{stiff, c->vel4, s->fluid.vel, nullptr, dt};
    return run_sweep(c, s, op);
}

extern "C" int sphk_pbd_density_lambda(sphk_ctx* c, const sphk_scene* s, float* lambda, float rho0, float relaxation) {
    SPHK_CHECK_SCENE(c, s);
    if (!lambda || !s->fluid.density) return SPHK_ERR_INVALID;
    OpPbdLambda op{s->fluid.density, lambda, rho0, (rho0 != 0.0f) ? 1.0f : 0.0f, relaxation};
    return run_sweep(c, s, op);
}

extern "C" int sphk_pbd_delta_pos_apply(sphk_ctx* c, const sphk_scene* s, const float* lambda, float* delta_pos,
                                        float rho0, const float space[3]) {
    SPHK_CHECK_SCENE(c, s);
    if (!lambda || !delta_pos || !space) return SPHK_ERR_INVALID;
    OpScalarGradient<2> op{lambda, nullptr, nullptr, delta_pos, rho0};
    const int rc = run_sweep(c, s, op);
    if (rc != SPHK_OK) return rc;
    k_apply_delta_pos<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(c->posm, s->fluid.pos, delta_pos, c->nF,
                                                                       make_float3(space[0], space[1], space[2]));
    c->launches++;
    c->posDirty = true;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_pbd_velocity_from_positions(sphk_ctx* c, const sphk_scene* s, const float* pos_last, float dt) {
    SPHK_CHECK_SCENE(c, s);
    if (!pos_last) return SPHK_ERR_INVALID;
    k_vel_from_pos<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(c->posm, pos_last, c->vel4, s->fluid.vel, c->nF, dt);
    c->launches++;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_pbd_xsph(sphk_ctx* c, const sphk_scene* s, float xc, float rho0) {
    SPHK_CHECK_SCENE(c, s);
    float4* tmp = c->snapB;
    OpXsph op{c->vel4, tmp, xc, rho0};
    const int rc = run_sweep(c, s, op);
    if (rc != SPHK_OK) return rc;
    k_commit_vel<<<sphk_blocks(c->nF), SPHK_BLOCK, 0, c->stream>>>(tmp, c->vel4, s->fluid.vel, c->nF);
    c->launches++;
    SPHK_CUDA_TRY(cudaGetLastError());
    return SPHK_OK;
}

extern "C" int sphk_list_stats(sphk_ctx* c, const sphk_scene* s, long long out_host[3]) {
    SPHK_CHECK_SCENE(c, s);
    if (!out_host) return SPHK_ERR_INVALID;
    DevScene d = dev_scene(c, s);
    const int rc = ensure_list(c, d);
    if (rc != SPHK_OK) return rc;
    unsigned long long* dev = reinterpret_cast<unsigned long long*>(c->partial);
    SPHK_CUDA_TRY(cudaMemsetAsync(dev, 0, 3 * sizeof(unsigned long long), c->stream));
    k_list_stats<<<256, 256, 0, c->stream>>>(c->cnt, c->nF, c->kmax, dev);
    c->launches++;
    unsigned long long h[3];
    SPHK_CUDA_TRY(cudaMemcpyAsync(h, dev, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    SPHK_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; ++k) out_host[k] = static_cast<long long>(h[k]);
    return SPHK_OK;
}
