"""tnrkit.jl_b200 -- B200-native engine for TNRKit's coarse-graining hot path.

Host-side mirror (Python; Julia is not available in this image) of the reference
interface for that path: scheme constructors, `run!`, `truncrank`, `maxiter`,
`classical_*` models, `free_energy`.  Everything numerical runs in `lib/libtnrcuda.so`
(hand-written CUDA for sm_100a) through the C ABI declared in include/tnrcuda.h."""
from . import _lib
from ._lib import Context, TNRCudaError, default_context
from .cft import (GSDegeneracy_Finalizer, central_charge, central_charge_Finalizer, cft_data,
                  finalize_central_charge, finalize_groundstatedegeneracy, finalize_gu_wen_ratio,
                  ground_state_degeneracy, gu_wen_ratio, guwenratio_Finalizer, transfer_matrix)
from .free_energy import free_energy
from .models import (XY_bc, XY_βc, ChargedArray, Trivial, U1Irrep, Z2Irrep, ZNIrrep, classical_XY,
                     classical_clock,
                     classical_ising, classical_ising_3D, classical_potts, f_onsager, ising_bc,
                     ising_bc_3D, ising_cft_exact, ising_βc, ising_βc_3D, phi4_complex, phi4_real, potts_bc, potts_βc,
                     sixvertex)
from .schemes import (ATRG, ATRG_3D, BTRG, HOTRG, HOTRG_3D, TRG, Finalizer, TNRScheme,
                      allgather_last_leg, beta_sweep, default_Finalizer, finalize,
                      finalize_two_by_two, run, run_, two_by_two_Finalizer,
                      shard_range)
from .stopping import MultipleCrit, convcrit, maxiter, stopcrit, trivial_convcrit
from .symmetric import Leg, SymTensor, sym_contract, sym_svd_trunc
from .tensor import DeviceTensor, contract, eigh_trunc, svd_trunc
from .truncation import truncrank, trunctol

__all__ = [n for n in dir() if not n.startswith("_")]
