// Scheme-level entry points (dense sector).  See schemes.cu.
#pragma once
#include "tensor.hpp"

namespace tnr {

double finalize_2d(Context* ctx, DT& T);
double finalize_btrg(Context* ctx, DT& T, const DT& S1, const DT& S2);
double finalize_3d(Context* ctx, DT& T);

DT trg_step(Context* ctx, const DT& T, int chi);
void btrg_step(Context* ctx, DT& T, DT& S1, DT& S2, double kexp, int chi);
DT hotrg_step(Context* ctx, const DT& T, int chi);
DT atrg_step(Context* ctx, const DT& T, int chi);

Dims hotrg3d_substep_dims(const Dims& d, int chi);
// peers != nullptr: the T' buffers of all `npeers` ranks (own buffer included, peer-mapped
// device pointers); every slab is stored to all of them by the producing kernel
void hotrg3d_substep(Context* ctx, const DT& T, int chi, DT& Tout, long long f0, long long f1,
                     double* const* peers, int npeers);
// the two halves of a z-compression, separately (multi-GPU: the four truncated
// eigendecompositions are dealt to different ranks and exchanged, then everybody contracts)
//   which = 0: x-bond from MM^dagger (left), 1: x-bond from M^dagger M (right), 2 / 3: y-bond
Trunc hotrg3d_proj_half(Context* ctx, const DT& T, int which, int chi);
// (eps_left > eps_right) ? U_right : U_left, decided on the device (hotrg3d.jl:96)
DT hotrg3d_pick(Context* ctx, const DT& Ul, const double* eps_l, const DT& Ur, const double* eps_r);
void hotrg3d_contract(Context* ctx, const DT& T, const DT& Ux, const DT& Uy, DT& Tout,
                      long long f0, long long f1, double* const* peers, int npeers);
DT hotrg3d_step(Context* ctx, const DT& T, int chi);
DT atrg3d_step(Context* ctx, const DT& T, int chi);

}  // namespace tnr
