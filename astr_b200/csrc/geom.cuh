// astr_b200/csrc/geom.cuh -- device-side gridgeom (src/geom.F90:99-700).
#pragma once
#include "common.cuh"
#include "pointwise.cuh"
#include "../../include/astr_gpu.h"

// helpers exported by api.cu
int astr_sweep_slots(int d, int optype, const int* in_slots, const int* out_slots, int nf, int epi, int o_lo,
                     int o_hi);
int astr_exchange_slots(const int* slots, int nf, int d, int mode);
double* astr_slot_ptr(int slot);
int astr_xhalo_exchange(int d);   // gridsendrecv of direction d (multi-block)

// x must be in slots S_G+0..2 (nodes 0..im,0..jm,0..km); fills S_JAC and S_DXI.
int geom_gridgeom(const Layout& L, const astr_cfg& cfg, cudaStream_t st);

// face kernels of gridsendrecv (src/parallel.F90:2780-3035), used by api.cu
int geom_xhalo_single(const Layout& L, double* x3[3], int d, cudaStream_t st);
int geom_xhalo_pack(const Layout& L, double* x3[3], int d, double* buf_lo, double* buf_hi, cudaStream_t st);
int geom_xhalo_unpack(const Layout& L, double* x3[3], int d, const double* from_lo, const double* from_hi,
                      cudaStream_t st);
