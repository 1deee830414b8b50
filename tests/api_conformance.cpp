// api_conformance.cpp -- compile-time statement of the class API the reference application is written against
// (the signatures below are read off /root/reference/src/{DArray,Particles,SPHParticles,BaseSolver,
// BasicSPHSolver,DFSPHSolver,PBDSolver,SPHSystem}.h).  This ONE file is compiled (syntax only) against
//   (1) this repository's headers  (cpp-fluid-particles_b200/host/),  by g++
//   (2) the reference's own headers (/root/reference/src + the helper_math.h stand-in), by nvcc  -- when present
// so every static_assert holds for both engines: the drop-in claim, member by member.  Nothing here runs.
#include <cstdio>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>
#ifdef SPH_APP_REFERENCE_ENGINE
#include <helper_math.h>
#include "global.h"
#endif
#include "DArray.h"
#include "Particles.h"
#include "SPHParticles.h"
#include "BaseSolver.h"
#include "BasicSPHSolver.h"
#include "DFSPHSolver.h"
#include "PBDSolver.h"
#include "SPHSystem.h"

using FluidsRef = std::shared_ptr<SPHParticles>&;
using BoundariesRef = const std::shared_ptr<SPHParticles>&;
using Cells = const DArray<int>&;
template <class A, class B> constexpr bool same = std::is_same<A, B>::value;

// ---- DArray<T>: RAII device array, explicit length constructor, non-copyable ---------------------------------------
static_assert(std::is_constructible<DArray<float3>, unsigned int>::value && !std::is_convertible<unsigned int, DArray<float3>>::value, "explicit DArray(length)");
static_assert(!std::is_copy_constructible<DArray<float>>::value && !std::is_copy_assignable<DArray<int>>::value, "DArray is non-copyable");
static_assert(same<decltype(&DArray<float3>::addr), float3* (DArray<float3>::*)(int) const>, "T* addr(int offset = 0) const");
static_assert(same<decltype(std::declval<const DArray<float>&>().addr()), float*>, "addr() has a default offset");
static_assert(same<decltype(&DArray<int>::length), unsigned int (DArray<int>::*)() const>, "unsigned length() const");
static_assert(same<decltype(&DArray<int>::clear), void (DArray<int>::*)()>, "void clear()");

// ---- Particles / SPHParticles ----------------------------------------------------------------------------------------
static_assert(std::is_constructible<Particles, const std::vector<float3>&>::value && !std::is_convertible<std::vector<float3>, Particles>::value, "explicit Particles(vector<float3>)");
static_assert(!std::is_copy_constructible<Particles>::value && !std::is_copy_assignable<Particles>::value, "Particles is non-copyable");
static_assert(std::has_virtual_destructor<Particles>::value, "virtual ~Particles");
static_assert(same<decltype(&Particles::size), unsigned int (Particles::*)() const>, "unsigned size() const");
static_assert(same<decltype(&Particles::getPosPtr), float3* (Particles::*)() const>, "float3* getPosPtr() const");
static_assert(same<decltype(&Particles::getVelPtr), float3* (Particles::*)() const>, "float3* getVelPtr() const");
static_assert(same<decltype(&Particles::getPos), const DArray<float3>& (Particles::*)() const>, "const DArray<float3>& getPos() const");
static_assert(same<decltype(&Particles::advect), void (Particles::*)(float)>, "void advect(float dt)");

static_assert(std::is_base_of<Particles, SPHParticles>::value && std::is_final<SPHParticles>::value, "SPHParticles final : Particles");
static_assert(std::is_constructible<SPHParticles, const std::vector<float3>&>::value && !std::is_copy_constructible<SPHParticles>::value, "explicit SPHParticles(vector<float3>)");
static_assert(same<decltype(&SPHParticles::getPressurePtr), float* (SPHParticles::*)() const>, "float* getPressurePtr() const");
static_assert(same<decltype(&SPHParticles::getDensityPtr), float* (SPHParticles::*)() const>, "float* getDensityPtr() const");
static_assert(same<decltype(&SPHParticles::getMassPtr), float* (SPHParticles::*)() const>, "float* getMassPtr() const");
static_assert(same<decltype(&SPHParticles::getParticle2Cell), int* (SPHParticles::*)() const>, "int* getParticle2Cell() const");
static_assert(same<decltype(&SPHParticles::getPressure), const DArray<float>& (SPHParticles::*)() const>, "const DArray<float>& getPressure() const");
static_assert(same<decltype(&SPHParticles::getDensity), const DArray<float>& (SPHParticles::*)() const>, "const DArray<float>& getDensity() const");

// ---- solvers: the 16-argument step, by value, in the reference's order ------------------------------------------------------
using StepFn = void(FluidsRef, BoundariesRef, Cells, Cells, float3 /*spaceSize*/, int3 /*cellSize*/, float /*cellLength*/,
                    float /*radius*/, float /*dt*/, float /*rho0*/, float /*rhoB*/, float /*stiff*/, float /*visc*/, float3 /*G*/,
                    float /*surfaceTensionIntensity*/, float /*airPressure*/);
template <class S> using StepPtr = void (S::*)(FluidsRef, BoundariesRef, Cells, Cells, float3, int3, float, float, float, float,
                                               float, float, float, float3, float, float);
static_assert(std::is_abstract<BaseSolver>::value && std::has_virtual_destructor<BaseSolver>::value, "BaseSolver is an interface");
static_assert(same<decltype(&BaseSolver::step), StepPtr<BaseSolver>>, "BaseSolver::step");
static_assert(same<decltype(&BasicSPHSolver::step), StepPtr<BasicSPHSolver>>, "BasicSPHSolver::step");
static_assert(same<decltype(&DFSPHSolver::step), StepPtr<DFSPHSolver>>, "DFSPHSolver::step");
static_assert(same<decltype(&PBDSolver::step), StepPtr<PBDSolver>>, "PBDSolver::step");
static_assert(std::is_base_of<BaseSolver, BasicSPHSolver>::value && !std::is_final<BasicSPHSolver>::value, "BasicSPHSolver : BaseSolver");
static_assert(std::is_base_of<BasicSPHSolver, DFSPHSolver>::value && std::is_final<DFSPHSolver>::value, "DFSPHSolver final : BasicSPHSolver");
static_assert(std::is_base_of<BasicSPHSolver, PBDSolver>::value && std::is_final<PBDSolver>::value, "PBDSolver final : BasicSPHSolver");
// constructors and their defaults (main.cpp:119-130 uses the one-argument forms)
static_assert(std::is_constructible<BasicSPHSolver, int>::value && !std::is_convertible<int, BasicSPHSolver>::value, "explicit BasicSPHSolver(int)");
static_assert(std::is_constructible<DFSPHSolver, int>::value && std::is_constructible<DFSPHSolver, int, float>::value &&
              std::is_constructible<DFSPHSolver, int, float, float>::value && std::is_constructible<DFSPHSolver, int, float, float, int>::value &&
              !std::is_convertible<int, DFSPHSolver>::value, "explicit DFSPHSolver(int, float = 1e-3f, float = 1e-3f, int = 20)");
static_assert(std::is_constructible<PBDSolver, int>::value && std::is_constructible<PBDSolver, int, int, float, float>::value &&
              !std::is_convertible<int, PBDSolver>::value, "explicit PBDSolver(int, int = 20, float = 0.05f, float = 0.75f)");
static_assert(std::is_constructible<PBDSolver, const std::shared_ptr<SPHParticles>&>::value &&
              std::is_constructible<PBDSolver, const std::shared_ptr<SPHParticles>&, int, float, float>::value, "explicit PBDSolver(particles, int = 20, float = 0.1f, float = 1.0f)");
static_assert(same<decltype(&PBDSolver::initializePosLast), void (PBDSolver::*)(const DArray<float3>&)>, "void initializePosLast(const DArray<float3>&)");
// the protected virtual hooks a user-defined solver overrides (BasicSPHSolver.h:31-43)
struct HookProbe : BasicSPHSolver {
    using BasicSPHSolver::BasicSPHSolver;
    static void signatures() {      // (inside the derived class: the hooks are protected)
        static_assert(same<decltype(&HookProbe::force), void (BasicSPHSolver::*)(FluidsRef, float, float3)>, "force(fluids, dt, G)");
        static_assert(same<decltype(&HookProbe::advect), void (BasicSPHSolver::*)(FluidsRef, float, float3)>, "advect(fluids, dt, spaceSize)");
        static_assert(same<decltype(&HookProbe::project), void (BasicSPHSolver::*)(FluidsRef, BoundariesRef, Cells, Cells, float, float, int3, float, float, float)>,
                      "project(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, stiff, cellSize, cellLength, radius, dt)");
        static_assert(same<decltype(&HookProbe::diffuse), void (BasicSPHSolver::*)(FluidsRef, Cells, int3, float, float, float, float, float)>,
                      "diffuse(fluids, cellStartFluid, cellSize, cellLength, rho0, radius, visc, dt)");
        static_assert(same<decltype(&HookProbe::handleSurface),
                           void (BasicSPHSolver::*)(FluidsRef, BoundariesRef, Cells, Cells, float, float, int3, float, float, float, float, float)>,
                      "handleSurface(fluids, boundaries, cellStartFluid, cellStartBoundary, rho0, rhoB, cellSize, cellLength, radius, dt, "
                      "surfaceTensionIntensity, airPressure)");
    }
};

// ---- SPHSystem: the 16-argument constructor of main.cpp:131-134, step() returning milliseconds, the accessors -----------------
static_assert(std::is_constructible<SPHSystem, std::shared_ptr<SPHParticles>&, std::shared_ptr<SPHParticles>&, std::shared_ptr<BaseSolver>&, float3, float,
                                    float, float, float, float, float, float, float, float, float, float3, int3>::value, "SPHSystem(16 arguments)");
static_assert(!std::is_constructible<SPHSystem, std::shared_ptr<SPHParticles>&&, std::shared_ptr<SPHParticles>&, std::shared_ptr<BaseSolver>&, float3, float,
                                     float, float, float, float, float, float, float, float, float, float3, int3>::value, "the particle handles are taken by non-const lvalue reference (the system moves from them)");
static_assert(!std::is_copy_constructible<SPHSystem>::value && !std::is_copy_assignable<SPHSystem>::value, "SPHSystem is non-copyable");
static_assert(same<decltype(&SPHSystem::step), float (SPHSystem::*)()>, "float step()");
static_assert(same<decltype(&SPHSystem::size), int (SPHSystem::*)() const>, "int size() const");
static_assert(same<decltype(&SPHSystem::fluidSize), int (SPHSystem::*)() const>, "int fluidSize() const");
static_assert(same<decltype(&SPHSystem::boundarySize), int (SPHSystem::*)() const>, "int boundarySize() const");
static_assert(same<decltype(&SPHSystem::totalSize), int (SPHSystem::*)() const>, "int totalSize() const");
static_assert(same<decltype(std::declval<const SPHSystem&>().getFluids()), std::shared_ptr<SPHParticles>>, "getFluids() returns a shared_ptr copy");
static_assert(same<decltype(std::declval<const SPHSystem&>().getBoundaries()), std::shared_ptr<SPHParticles>>, "getBoundaries() returns a shared_ptr copy");

// ---- the render hook, vbo.cu:46 (declared by the application, main.cpp:268) ---------------------------------------------------
extern "C" void generate_dots(float3* dot, float3* color, const std::shared_ptr<SPHParticles> particles);

int api_conformance_anchor() { return 0; }
