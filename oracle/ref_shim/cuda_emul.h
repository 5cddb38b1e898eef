/*
 * cuda_emul.h -- a minimal CUDA execution-model emulator for the CPU (TEST INFRASTRUCTURE, part of the oracle).
 *
 * Purpose: run the reference's own __global__ kernels (src/phdfilter.cu, src/device_math.cuh of
 * cheesinglee/cuda-PHDSLAM) unmodified on the host, so the CPU restatement in phd_oracle.cpp can be pinned
 * against the reference's actual arithmetic (oracle/ref_build.sh, oracle/ref_harness.cpp).  Nothing here is
 * product code and nothing here is taken from the reference.
 *
 * Model: one thread block at a time; every CUDA thread of the block is a ucontext fiber.  Fibers run in
 * ascending thread order until they reach a barrier:
 *   __syncthreads()  releases when every live thread of the block waits at a __syncthreads();
 *   __syncwarp()     releases when every live thread of the warp waits at a barrier, and takes priority.
 * Between two barriers the threads of a warp therefore execute one after the other in ascending lane order.
 * For the reference's warp-synchronous shared-memory reductions (sdata[tid] += sdata[tid+k], k = 32..1, one
 * __syncwarp() per step, inserted by ref_build.sh as any post-Volta port must) that order reproduces the
 * lock-step semantics exactly: lane t reads sdata[t+k] before lane t+k overwrites it.
 * atomicAdd is a plain read-modify-write (single host thread).
 */
#ifndef PHD_CUDA_EMUL_H
#define PHD_CUDA_EMUL_H

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __constant__
#define __shared__ static
#define __forceinline__ inline
#define __restrict__

struct emul_dim3 {
  unsigned x, y, z;
  emul_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef emul_dim3 dim3;
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }

extern emul_dim3 threadIdx, blockIdx, blockDim, gridDim;
void __syncthreads();
void __syncwarp();

template <class T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }

/* curand types only appear in signatures of functions the oracle pin never calls */
struct curandStateMRG32k3a { int unused; };
typedef curandStateMRG32k3a curandStateMRG32k3a_t;
static inline float curand_normal(curandStateMRG32k3a*) { abort(); }
static inline float curand_uniform(curandStateMRG32k3a*) { abort(); }
static inline float2 curand_normal2(curandStateMRG32k3a*) { abort(); }

/* run body() once per CUDA thread of a grid x block launch (1-D), with the barrier semantics above */
void emul_launch(unsigned grid, unsigned block, const std::function<void()>& body);

#ifdef PHD_CUDA_EMUL_IMPL
emul_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emul {
enum State { RUN, AT_BLOCK, AT_WARP, DONE };
struct Fiber { ucontext_t ctx; State st; char* stack; };
static const size_t kStack = 256 * 1024;
static std::vector<Fiber> fibers;
static std::vector<char*> stacks;
static ucontext_t sched_ctx;
static int current = -1;
static const std::function<void()>* body_ptr = nullptr;

static void trampoline() {
  (*body_ptr)();
  fibers[current].st = DONE;
  swapcontext(&fibers[current].ctx, &sched_ctx);
}
static void yield_as(State s) {
  if (current < 0) return;          /* called outside a launch (host code): no-op */
  int me = current;
  fibers[me].st = s;
  swapcontext(&fibers[me].ctx, &sched_ctx);
}
}  // namespace emul

void __syncthreads() { emul::yield_as(emul::AT_BLOCK); }
void __syncwarp() { emul::yield_as(emul::AT_WARP); }

void emul_launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  using namespace emul;
  gridDim = emul_dim3(grid);
  blockDim = emul_dim3(block);
  body_ptr = &body;
  while (stacks.size() < block) stacks.push_back((char*)malloc(kStack));
  fibers.resize(block);
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx = emul_dim3(b);
    for (unsigned t = 0; t < block; ++t) {
      Fiber& f = fibers[t];
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = stacks[t];
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = &sched_ctx;
      makecontext(&f.ctx, trampoline, 0);
      f.st = RUN;
    }
    for (;;) {
      bool any = false;
      for (unsigned t = 0; t < block; ++t) {
        if (fibers[t].st != RUN) continue;
        any = true;
        current = (int)t;
        threadIdx = emul_dim3(t);
        swapcontext(&sched_ctx, &fibers[t].ctx);
      }
      current = -1;
      if (any) continue;   /* re-scan: nobody is runnable now */
      /* warp barriers first */
      bool released = false;
      for (unsigned w = 0; w * 32 < block; ++w) {
        bool has = false;
        for (unsigned t = w * 32; t < std::min(block, w * 32 + 32); ++t) has |= fibers[t].st == AT_WARP;
        if (!has) continue;
        for (unsigned t = w * 32; t < std::min(block, w * 32 + 32); ++t)
          if (fibers[t].st == AT_WARP) fibers[t].st = RUN;
        released = true;
      }
      if (released) continue;
      bool all_done = true;
      for (unsigned t = 0; t < block; ++t) {
        if (fibers[t].st == AT_BLOCK) { fibers[t].st = RUN; all_done = false; }
      }
      if (all_done) break;
    }
  }
  body_ptr = nullptr;
}
#endif /* PHD_CUDA_EMUL_IMPL */

#endif
