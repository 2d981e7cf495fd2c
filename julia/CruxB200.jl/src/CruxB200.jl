"""
    CruxB200

Drop-in acceleration of the Crux.jl actor-learner hot path on NVIDIA B200 (sm_100a): `solve(::OnPolicySolver, envs)` and
`solve(::OffPolicySolver, envs)` for a VECTOR of POMDPs.jl models (one env stream each), with the policy forward, the replay /
rollout buffer, GAE and the whole update running in `libcrux_cuda.so` (C ABI: `include/crux_cuda.h`).  Everything above the hot path
stays Crux.jl's own code: the solver constructors (`PPO`, `A2C`, `DQN`, `SAC`, ...), `TrainingParams`, `LoggerParams`, callbacks.

    using Crux, CruxB200, POMDPs
    𝒮 = PPO(π=ActorCritic(GaussianPolicy(ContinuousNetwork(Chain(Dense(17,64,tanh), Dense(64,64,tanh), Dense(64,6))), -0.5f0*ones(Float32,6)),
                          ContinuousNetwork(Chain(Dense(17,64,tanh), Dense(64,64,tanh), Dense(64,1)))),
            S=ContinuousSpace(17), N=10*131072, ΔN=131072, a_opt=(epochs=4, batch_size=32768), c_opt=(epochs=4, batch_size=32768))
    solve(𝒮, [LinQuadMDP() for _ in 1:4096])          # Vector{<:MDP}  => this package; a single mdp => stock Crux.jl

File map (reference file:line each part replaces):
  abi.jl        library loading, status -> exception, C-layout structs               (include/crux_cuda.h)
  networks.jl   Chain(Dense...) / GaussianPolicy / DiscreteNetwork mirrors            (src/policies.jl:68-157,315-400)
  buffer.jl     ExperienceBuffer on the device                                         (src/experience_buffer.jl)
  sampler.jl    steps! over N env streams, fill_gae! / fill_returns!                   (src/sampler.jl:71-173,255-281)
  onpolicy.jl   solve(::OnPolicySolver, envs), policy_gradient_training                (src/model_free/on_policy.jl:56-109)
  offpolicy.jl  solve(::OffPolicySolver, envs), value_training for DQN / SAC           (src/model_free/off_policy.jl:66-150)
  linquad.jl    the synthetic LinQuad-17x6 MDP of the benchmark as a POMDPs.MDP + its device-resident twin
"""
module CruxB200

using Crux, Flux, CUDA, POMDPs, Libdl, Random
import POMDPs: solve

include("abi.jl")
include("networks.jl")
include("buffer.jl")
include("sampler.jl")
include("onpolicy.jl")
include("offpolicy.jl")
include("linquad.jl")

export Ctx, DevMLP, DevGaussian, DevCategorical, DevBuffer, VecSampler, LinQuadMDP, DeviceLinQuad, mirror, pull!, fill_gae_returns!, whiten!

end # module
