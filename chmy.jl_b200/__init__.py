"""
chmy_b200 -- B200-native hot path of PTsolvers/Chmy.jl behind the reference's own API names.

Julia `f!` is spelled `f_` here; `op => args` is the tuple `(op, args)`; 1-based Dim/Side are kept.
Everything computes through libchmy_b200.so (hand-written CUDA for sm_100a, include/chmy_b200.h).
"""
from ._lib import ChmyError, LIB_PATH, lib as load_library
from .utils import Dim, Side, Left, Right, remove_dim, insert_dim, DoubleBuffer, swap_, front, back
from .architectures import (Architecture, SingleDeviceArchitecture, DistributedArchitecture, B200Backend, Arch,
                            get_backend, get_device, activate_, set_device_, is_gpu_aware, synchronize, launch_count, topology,
                            event_record, event_elapsed_ms, time_fused_sweep, set_fusion, fused_count, fusion_fallback_count, last_division_mode, division_two_op_exact, overlapped_count, set_fused_tuning, set_fused2d_tuning, set_launch_split,
                            set_exchange_mode, exchange_stats)
from .grids import (Location, Center, Vertex, flip, Bounded, Connected, UniformAxis, StructuredGrid, UniformGrid,
                    connectivity, spacing, inv_spacing, coord, coords, centers, vertices, origin, extent, bounds,
                    axes_names, expand_loc, nvertices, ncenters, axis, vertex, center, direction, volume, inv_volume)
from .fields import (AbstractField, ConstantField, ZeroField, OneField, ValueField, Field, FieldTuple, VectorField, TensorField, FunctionField, init_incl, init_gauss, set_,
                     interior, parent, fill_parent_, halo, location, maxabs, maxabs_many, vector_location, pinned_array)
from .boundary_conditions import (BoundaryFunction, FirstOrderBC, Dirichlet, Neumann, EmptyBatch, FieldBatch, ExchangeBatch, batch, bc_)
from .kernel_launch import (Launcher, worksize, outer_width, inner_worksize, inner_offset, outer_worksize,
                            outer_offset)
from .distributed import (CartesianTopology, TorchDistComm, dims_create, exchange_halo_, allreduce_max, barrier, gather_,
                          global_rank, shared_rank, node_name, dims, cart_coords, neighbors, neighbor, has_neighbor,
                          global_size, node_size, cart_comm, shared_comm, PROC_NULL)
from .ops import (KernelOp, compute_q_, update_C_, update_old_, update_stress_, update_velocity_,
                  update_thermal_flux_, update_thermal_)
from .grid_operators import (left_, right_, delta_, partial_, partial2_, dkd_, dx_, dy_, dz_, d2x_, d2y_, d2z_, lerp_, hlerp_,
                             divg_, lapl_, divg_grad_, vmag_, grad_, kgrad_)
