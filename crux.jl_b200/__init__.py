"""crux.jl_b200 -- host-side mirror of the Crux.jl solver surface driving libcrux_cuda.so (sm_100a).

Only what the hot path needs: policies, ExperienceBuffer, Sampler, PPO/A2C/DQN/SAC + solve.
Import as ``import crux_b200 as crux`` (see crux_b200.py at the repo root).  There is no CPU fallback:
constructing anything that touches the device raises if the CUDA library or a GPU is missing.
"""
from . import _abi  # noqa: F401
from ._abi import CruxError, NaNError  # noqa: F401
from .device import Context, default_context  # noqa: F401
from .spaces import ContinuousSpace, DiscreteSpace, dim, state_space, tovec, whiten  # noqa: F401
from .policies import (ActorCritic, Chain, ContinuousNetwork, Conv, Dense, DiscreteNetwork, DoubleNetwork,  # noqa: F401
                       FirstExplorePolicy, GaussianNoiseExplorationPolicy, GaussianPolicy, LinearDecaySchedule,
                       MixedPolicy, PolicyParams, SquashedGaussianPolicy, action, action_space, actor, copyto_,
                       critic, deepcopy, entropy, eps_greedy_policy, exploration, flatten, glorot_uniform, identity, logpdf,
                       polyak_average_, relu, scale255, tanh, value)
from .buffer import (ExperienceBuffer, PriorityParams, buffer_like, mdp_data, prioritized_sample_, rand_, split_batches,  # noqa: F401
                     uniform_sample_)
from .envs import DeviceLinQuad, HostLinQuad, NativeHostLinQuad, SimpleGridWorld, linquad_matrices  # noqa: F401
from .sampler import Sampler, fill_gae_, fill_returns_, steps_  # noqa: F401
from .logger import (LoggerParams, TBLogger, aggregate_info, log_discounted_return, log_episode_averages, log_experience_sums,  # noqa: F401
                     log_exploration, log_failure, log_metric_by_key, log_metrics_by_key, log_performance, log_undiscounted_return,
                     log_validation_error, read_scalars, tb_increment)
from .solvers import (A2C, DDPG, DQN, PPO, LagrangePPO, REINFORCE, SAC, TD3, Adam, SoftQ, OffPolicySolver, OnPolicySolver, TrainingParams,  # noqa: F401
                      solve)
