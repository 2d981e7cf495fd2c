# ---- the synthetic benchmark MDP "LinQuad-17x6" (SURVEY 8d) as a POMDPs.jl model, and its device-resident twin --------------------------------------
# s' = clip(A s + B tanh(a) + 0.01 ξ, -10, 10), A = 0.95 I + 0.02 G1, B = 0.1 G2; r = 1 - |s'|²/S - 0.1 |a|²/A; terminal if |s'_1| > 5;
# s0 ~ U(-0.1, 0.1)^S; γ = 0.99.  (G1, G2 are fixed matrices shared by every copy; the Python side draws them from numpy's default_rng(0),
# here they are passed in or drawn once from a seeded Julia RNG -- the benchmark numbers do not depend on the particular matrices.)
struct LinQuadMDP <: MDP{Vector{Float32},Vector{Float32}}
    A::Matrix{Float32}
    B::Matrix{Float32}
    γ::Float32
end
function LinQuadMDP(; sdim::Int=17, adim::Int=6, γ=0.99f0, rng=Random.MersenneTwister(0))
    A = Float32[i == j ? 0.95f0 : 0f0 for i in 1:sdim, j in 1:sdim] .+ 0.02f0 .* randn(rng, Float32, sdim, sdim)
    LinQuadMDP(A, 0.1f0 .* randn(rng, Float32, sdim, adim), Float32(γ))
end
POMDPs.discount(m::LinQuadMDP) = m.γ
"s0 ~ U(-0.1, 0.1)^S as a distribution object: `rand(initialstate(m))` (sampler.jl:44) and `rand(rng, initialstate(m))` both work"
struct LinQuadInit
    n::Int
end
Base.eltype(::Type{LinQuadInit}) = Vector{Float32}
Random.rand(rng::Random.AbstractRNG, d::Random.SamplerTrivial{LinQuadInit}) = (rand(rng, Float32, d[].n) .* 2f0 .- 1f0) .* 0.1f0
POMDPs.initialstate(m::LinQuadMDP) = LinQuadInit(size(m.A, 1))
POMDPs.isterminal(m::LinQuadMDP, s) = abs(s[1]) > 5f0
POMDPs.convert_s(::Type{<:AbstractArray}, s::Vector{Float32}, ::LinQuadMDP) = s
POMDPs.actions(m::LinQuadMDP) = Crux.ContinuousSpace(size(m.B, 2))
function POMDPs.gen(m::LinQuadMDP, s, a, rng=Random.default_rng())
    sp = clamp.(m.A * s .+ m.B * tanh.(a) .+ 0.01f0 .* randn(rng, Float32, length(s)), -10f0, 10f0)
    r = 1f0 - sum(abs2, sp) / length(sp) - 0.1f0 * sum(abs2, a) / length(a)
    (sp=sp, r=r)
end

"the same MDP stepped on the device (`crux_linquad_*`): N streams, rollouts of T vector steps in ONE persistent launch"
mutable struct DeviceLinQuad
    h::Ptr{Cvoid}
    N::Int
    sdim::Int
    adim::Int
    max_steps::Int
    γ::Float32
    obs::CuArray{Float32,2}           # current observation of every stream
end
function DeviceLinQuad(m::LinQuadMDP, N::Int; max_steps::Int=1000, seed::Integer=0)
    sd, ad = size(m.B)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    # the ABI takes row-major [sdim][sdim] / [sdim][adim]: the transpose of Julia's column-major matrices
    chk(ccall(sym(:crux_linquad_create), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float32}, Ptr{Float32}, Int64, Int32, UInt64, Ref{Ptr{Cvoid}}),
              ctx().h, sd, ad, collect(m.A'), collect(m.B'), N, max_steps, seed, out), ctx().h)
    env = DeviceLinQuad(out[], N, sd, ad, max_steps, m.γ, CUDA.zeros(Float32, sd, N))
    chk(ccall(sym(:crux_linquad_reset), Int32, (Ptr{Cvoid}, CuPtr{Float32}), env.h, env.obs), ctx().h)
    finalizer(e -> (e.h != C_NULL && ccall(sym(:crux_linquad_destroy), Int32, (Ptr{Cvoid},), e.h); e.h = C_NULL), env)
end
POMDPs.discount(e::DeviceLinQuad) = e.γ

"`steps!` for the device env: policy forward, Gaussian sample, transition, bookkeeping and resets of T vector steps in one launch (`crux_linquad_rollout`)"
function Crux.steps!(env::DeviceLinQuad, π::DevGaussian, 𝒟::DevBuffer; Nsteps::Int, reset::Bool=true, seed::Integer=0, ctr::Integer=0)
    T = Nsteps ÷ env.N
    cols = Ref(RolloutCols(pointer(𝒟.cols[:s]), pointer(𝒟.cols[:a]), pointer(𝒟.cols[:sp]), pointer(𝒟.cols[:r]), pointer(𝒟.cols[:done]),
                           pointer(𝒟.cols[:episode_end]), haskey(𝒟, :logprob) ? pointer(𝒟.cols[:logprob]) : CU_NULL))
    chk(ccall(sym(:crux_linquad_rollout), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, CuPtr{Float32}, Ref{RolloutCols}, UInt64, UInt64),
              env.h, π.h, T, reset ? 1 : 0, env.obs, cols, seed, ctr), ctx().h)
    commit_rows!(𝒟, Nsteps)
    T
end

"`solve(𝒮::OnPolicySolver, env::DeviceLinQuad)`: the on-policy loop with the environment on the GPU (the `value` leg of bench.py)"
function solve(𝒮::OnPolicySolver, env::DeviceLinQuad)
    π = 𝒮.agent.π
    g = mirror(π.A); V = 𝒮.c_opt !== nothing ? mirror(π.C) : nothing
    set_adam!(g.mu, 𝒮.a_opt.optimizer); V !== nothing && set_adam!(V, 𝒮.c_opt.optimizer)
    𝒮.max_steps == env.max_steps || error("a device env bakes max_steps into its step kernel: construct it with the solver's max_steps")
    T = 𝒮.ΔN ÷ env.N
    𝒟 = DevBuffer(𝒮.S, 𝒮.agent.space, 𝒮.ΔN, Symbol.(𝒮.required_columns))
    ctr = 0
    for 𝒮.i in range(𝒮.i, stop=𝒮.i + 𝒮.N - 𝒮.ΔN, step=𝒮.ΔN)
        info = Dict()
        clear!(𝒟)
        steps!(env, g, 𝒟; Nsteps=𝒮.ΔN, reset=true, ctr=ctr); ctr += T
        fill_gae_returns!(𝒟, V, T, env.N, env.γ, 𝒮.λ_gae)
        𝒮.post_sample_callback(𝒟, info=info, 𝒮=𝒮)
        𝒮.post_batch_callback(𝒟, info=info, 𝒮=𝒮)
        policy_gradient_training(𝒮, g, V, 𝒟)
    end
    𝒮.i += 𝒮.ΔN
    pull!(g); V !== nothing && pull!(V)
    check_flags()
    π
end
