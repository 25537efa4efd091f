"""gaudi_b200 -- B200-native (sm_100a) guided-sampling hot path of GaUDI behind GaUDI's own Python surface.

Public surface (same names / signatures as the reference):
  EnVariationalDiffusion, EGNN_dynamics, EGNN, EquivariantBlock, GCL, EquivariantUpdate   (denoiser)
  EGNN_predictor, E_GCL                                                                   (property predictor)
  sample_guidance, sample_pos_edm, node2edge_mask, get_model, get_cond_predictor_model     (helpers)
plus AffineTarget (fused-loop cond_fn), args_edm / prediction_args presets and gaudi_b200.dist (batch sharding).
Training step: EnVariationalDiffusion.forward in train mode (gaudi_b200.training).  Geometric validity of the generated
ring graphs: gaudi_b200.analyze (positions2adj, check_stability, analyze_validity_for_molecules, check_stability_batch).
"""
from .diffusion import AffineTarget, EnVariationalDiffusion, PredefinedNoiseSchedule
from .egnn import EGNN, EGNN_dynamics, EquivariantBlock, EquivariantUpdate, GCL
from .predictor import E_GCL, EGNN_predictor
from .sampling import (DistributionProperty, DistributionRings, MyDataParallel, args_edm, build_masks,
                       get_cond_predictor_model, get_model, load_state_dict_flexible, node2edge_mask,
                       prediction_args, sample_guidance, sample_pos_edm, switch_grad_off)

from . import analyze, training  # noqa: E402,F401
from .analyze import (analyze_validity_for_molecules, check_stability, check_stability_batch,  # noqa: E402,F401
                      eval_geometric_stability, positions2adj)

__version__ = "0.1.0"
