"""tamf_b200 -- B200-native drop-in for the reverse-diffusion sampling hot path of OakInk2-TaMF.

Same call signatures as the reference (InterationSegmentMDM.forward(x, timesteps, batch), p_sample_loop,
ManoLayer(...)(pose_coeffs, betas), ChamferDistance()(x, y), point2point_signed, SegmentRefineModel(mano_path, ...)(batch)), backed by libtamf_b200.so."""
from .chamfer import ChamferDistance, H2OIndex, h2o_dist, nn_query, point2point_signed  # noqa: F401
from .diffusion import (GaussianDiffusion, SpacedDiffusion, create_gaussian_diffusion,  # noqa: F401
                        space_timesteps)
from .extract_sample import (contact_min_cdist, extract_refined_sample, extract_refined_sample_bihand,  # noqa: F401
                             extract_refined_samples, interaction_segment_collate, map_copy_select_to,
                             refine_dataset, sample_dataset, transf_merge_obj_pointcloud)
from .manolayer import MANOOutput, ManoLayer  # noqa: F401
from .mdm import InterationSegmentMDM  # noqa: F401
from .refine import SegmentRefineModel, vertex_normals  # noqa: F401
