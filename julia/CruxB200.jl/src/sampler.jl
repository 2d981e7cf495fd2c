# ---- steps! over N env streams (src/sampler.jl:71-173) -----------------------------------------------------------------------------------------
# One VecSampler drives N independent copies of the user's POMDPs.jl model (stream e is one reference `Sampler`).  Rollout rows are
# laid out [T][N] (column t*N + e of a `[features, ΔN]` array: the reference's own interleaving `j = (step-1)*Nenv + env`,
# sampler.jl:157-165); every stream follows the single-env rules of `step!`: episode_length, `done || episode_length >= max_steps`
# -> episode_end + reset, and the forced termination of `steps!(reset=true)` (:148).
mutable struct VecSampler{M<:MDP}
    mdps::Vector{M}
    states::Vector{Any}                 # POMDPs state of every stream
    obs::Matrix{Float32}                # pinned [sdim, N]: current (tovec'ed) observation of every stream
    episode_length::Vector{Int32}
    S::Crux.AbstractSpace
    max_steps::Int
    γ::Float32
    λ::Float32
    noise_ctr::UInt64                   # Philox stream position of the exploration noise
    seed::UInt64
end
svec(mdp, s, S) = Float32.(vec(Crux.tovec(convert_s(AbstractArray, s, mdp), S)))
function VecSampler(mdps::Vector{<:MDP}, S::Crux.AbstractSpace; max_steps::Int=100, λ=NaN32, seed::Integer=0)
    N, sd = length(mdps), prod(Crux.dim(S))
    states = Any[rand(initialstate(m)) for m in mdps]
    obs = pinned(Float32, sd, N)
    for e in 1:N
        obs[:, e] .= svec(mdps[e], states[e], S)
    end
    VecSampler(mdps, states, obs, zeros(Int32, N), S, max_steps, Float32(discount(mdps[1])), Float32(λ), UInt64(0), UInt64(seed))
end
"`reset_sampler!` (sampler.jl:31-43) for every stream"
function reset!(s::VecSampler)
    for e in eachindex(s.mdps)
        s.states[e] = rand(initialstate(s.mdps[e]))
        s.obs[:, e] .= svec(s.mdps[e], s.states[e], s.S)
    end
    fill!(s.episode_length, 0)
    s
end

# ---- the two callbacks crux_rollout_host reaches the environment through (include/crux_cuda.h: crux_env_step_fn / crux_env_reset_fn) ----------------
# `user` is a pointer to a Ref{Any} holding (sampler, adim); arrays are the library's pinned staging buffers.
function env_step_cb(user::Ptr{Cvoid}, e0::Int32, e1::Int32, a::Ptr{Float32}, sp::Ptr{Float32}, r::Ptr{Float32}, done::Ptr{UInt8})::Cvoid
    s, adim = unsafe_pointer_to_objref(user)[]::Tuple{VecSampler,Int}
    sd = size(s.obs, 1)
    for e in (e0 + 1):e1                                             # streams [e0, e1) 0-based
        act = unsafe_wrap(Array, a + (e - 1) * adim * sizeof(Float32), adim)
        act1 = adim == 1 ? act[1] : copy(act)                        # sampler.jl:74: length-1 actions are passed as scalars
        spe, re = @gen(:sp, :r)(s.mdps[e], s.states[e], act1)        # sampler.jl:92
        s.states[e] = spe
        unsafe_wrap(Array, sp + (e - 1) * sd * sizeof(Float32), sd) .= svec(s.mdps[e], spe, s.S)
        unsafe_store!(r, Float32(re), e)
        unsafe_store!(done, isterminal(s.mdps[e], spe) ? 0x01 : 0x00, e)
    end
    nothing
end
function env_reset_cb(user::Ptr{Cvoid}, idx::Ptr{Int32}, n::Int32, obs_out::Ptr{Float32})::Cvoid
    s, _ = unsafe_pointer_to_objref(user)[]::Tuple{VecSampler,Int}
    sd = size(s.obs, 1)
    for q in 1:n
        e = unsafe_load(idx, q) + 1
        s.states[e] = rand(initialstate(s.mdps[e]))                  # sampler.jl:39
        unsafe_wrap(Array, obs_out + (q - 1) * sd * sizeof(Float32), sd) .= svec(s.mdps[e], s.states[e], s.S)
    end
    nothing
end

"""
    steps!(s::VecSampler, π::DevGaussian, 𝒟::DevBuffer; Nsteps, reset=true)

`steps!(sampler, buffer; Nsteps, explore=true, reset)` (sampler.jl:139-155) for a Gaussian policy: the whole loop of `Nsteps ÷ N` vector
steps in ONE library call (`crux_rollout_host`); the rows are written in place into `𝒟` (its capacity is ΔN for on-policy solvers).
"""
function Crux.steps!(s::VecSampler, π::DevGaussian, 𝒟::DevBuffer; Nsteps::Int, reset::Bool=true)
    N = length(s.mdps)
    Nsteps % N == 0 || error("Nsteps=$Nsteps must be a multiple of the $N env streams")
    T = Nsteps ÷ N
    start = state(𝒟).next_ind
    start + Nsteps - 1 <= 𝒟.capacity || error("the rollout must fit the buffer without wrapping (capacity = ΔN for on-policy solvers)")
    col(k) = pointer(𝒟.cols[k], (start - 1) * size(𝒟.cols[k], 1) + 1)
    cols = Ref(RolloutCols(col(:s), col(:a), col(:sp), col(:r), col(:done), col(:episode_end), haskey(𝒟, :logprob) ? col(:logprob) : CU_NULL))
    user = Ref{Any}((s, π.adim))
    step_c = @cfunction(env_step_cb, Cvoid, (Ptr{Cvoid}, Int32, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}))
    reset_c = @cfunction(env_reset_cb, Cvoid, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ptr{Float32}))
    GC.@preserve user cols s begin
        chk(ccall(sym(:crux_rollout_host), Int32,
                  (Ptr{Cvoid}, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ref{RolloutCols}, UInt64, UInt64),
                  π.h, N, T, s.max_steps, reset ? 1 : 0, step_c, reset_c, pointer_from_objref(user), s.obs, s.episode_length, cols, s.seed, s.noise_ctr), ctx().h)
    end
    s.noise_ctr += T
    commit_rows!(𝒟, Nsteps)
    T
end

"""
    host_loop!(s, 𝒟, Nsteps, act; reset=false)

The generic `steps!` loop for policies without a one-call rollout: per vector step ONE upload of the observations, `act(s_dev, t)` ->
`(stored action rows [adim, N] on the device, per-stream env actions on the host[, extra columns as a Dict])`, the envs stepped on the host,
the rows pushed to `𝒟`.  `reset=true`: every stream's episode is terminated at the last vector step (`steps!(reset=true)`, sampler.jl:148-151).
"""
function host_loop!(s::VecSampler, 𝒟::DevBuffer, Nsteps::Int, act; reset::Bool=false)
    N, sd = length(s.mdps), size(s.obs, 1)
    T = cld(Nsteps, N)
    sp, r = zeros(Float32, sd, N), zeros(Float32, 1, N)
    done, ee = zeros(UInt8, 1, N), zeros(UInt8, 1, N)
    for t in 1:T
        sdev = CuArray(s.obs)                                                     # H2D: svec of every stream
        res = act(sdev, t)                                                        # D2H inside: the env needs the action
        a_rows, a_env = res[1], res[2]
        sobs = copy(s.obs)
        for e in 1:N
            spe, re = @gen(:sp, :r)(s.mdps[e], s.states[e], a_env[e])
            s.states[e] = spe
            sp[:, e] .= svec(s.mdps[e], spe, s.S); r[1, e] = re
            done[1, e] = isterminal(s.mdps[e], spe)
            s.episode_length[e] += 1                                              # sampler.jl:130
            fin = done[1, e] != 0 || s.episode_length[e] >= s.max_steps || (reset && t == T)
            ee[1, e] = fin
            if fin                                                                # terminate_episode! -> reset_sampler!
                s.states[e] = rand(initialstate(s.mdps[e])); s.episode_length[e] = 0
                s.obs[:, e] .= svec(s.mdps[e], s.states[e], s.S)
            else
                s.obs[:, e] .= sp[:, e]
            end
        end
        row = Dict{Symbol,Any}(:s => sobs, :a => Array(a_rows), :sp => sp, :r => r, :done => done, :episode_end => ee)
        length(res) >= 3 && merge!(row, res[3])
        push!(𝒟, row)
    end
    T * N
end

"on-policy rollouts of a categorical actor (examples/rl/cartpole.jl): one batched forward + `crux_discrete_explore` per vector step, `:logprob` stored"
function Crux.steps!(s::VecSampler, c::DevCategorical, 𝒟::DevBuffer; Nsteps::Int, reset::Bool=true)
    host_loop!(s, 𝒟, Nsteps, (sdev, t) -> begin
        idx, oh, lp = exploration(c, sdev; seed=s.seed, ctr=s.noise_ctr)
        s.noise_ctr += 1
        oh, [c.d.outputs[k] for k in Array(idx)], Dict{Symbol,Any}(:logprob => Array(lp))
    end; reset=reset)
end

"DQN: ϵ-greedy exploration of a DiscreteNetwork (policies.jl:474-494): one batched forward + `crux_discrete_eps_greedy` per vector step"
function Crux.steps!(s::VecSampler, q::DevDiscrete, ϵ::Function, 𝒟::DevBuffer; Nsteps::Int, i::Int=0)
    N = length(s.mdps)
    host_loop!(s, 𝒟, Nsteps, (sdev, t) -> begin
        idx, oh, _ = eps_greedy(q, sdev, ϵ(i + (t - 1) * N); seed=s.seed, ctr=s.noise_ctr)
        s.noise_ctr += 1
        oh, [q.outputs[k] for k in Array(idx)]
    end)
end

"SAC / DDPG / TD3: `exploration(::GaussianNoiseExplorationPolicy)` policies.jl:510-514 = clamp(action(π, s) + clamp(σ(i)·ε, ϵ_min, ϵ_max), a_min, a_max)"
function Crux.steps!(s::VecSampler, g::DevGaussian, ne::GaussianNoiseExplorationPolicy, 𝒟::DevBuffer; Nsteps::Int, i::Int=0)
    N = length(s.mdps)
    host_loop!(s, 𝒟, Nsteps, (sdev, t) -> begin
        a = action(g, sdev)
        chk(ccall(sym(:crux_noise_explore), Int32,
                  (Ptr{Cvoid}, CuPtr{Float32}, Int64, Int32, Float32, Float32, Float32, Ptr{Float32}, Int32, Ptr{Float32}, Int32, CuPtr{Float32}, UInt64, UInt64),
                  ctx().h, a, N, g.adim, Float32(ne.σ(i + (t - 1) * N)), ne.ϵ_min, ne.ϵ_max, ne.a_min, all(isinf, ne.a_min) ? 0 : length(ne.a_min),
                  ne.a_max, all(isinf, ne.a_max) ? 0 : length(ne.a_max), CU_NULL, s.seed, s.noise_ctr), ctx().h)
        s.noise_ctr += 1
        ah = Array(a)
        a, [g.adim == 1 ? ah[1, e] : ah[:, e] for e in 1:N]
    end)
end

# ---- fill_gae! + fill_returns! (sampler.jl:255-281) for every episode range of every stream: ONE segmented scan ---------------------------------------
"""
    fill_gae_returns!(𝒟, V, T, N, γ, λ)

`value(V, s)` over the whole rollout, `value(V, sp)` with V(s)[t+1] reused wherever sp[t] is bitwise s[t+1] (`crux_value_next`), then
`A = λγ·A + r + (1-done)·γ·V(sp) - V(s)` cut at `episode_end` and `R = r + γR` (no bootstrap) written into `:advantage` / `:return`.
"""
function fill_gae_returns!(𝒟::DevBuffer, V::Union{DevMLP,Nothing}, T::Int, N::Int, γ::Float32, λ::Float32)
    n = T * N
    adv = haskey(𝒟, :advantage) ? pointer(𝒟.cols[:advantage]) : CU_NULL
    ret = haskey(𝒟, :return) ? pointer(𝒟.cols[:return]) : CU_NULL
    v_s = v_sp = 𝒟.cols[:r]                                                        # unused by the kernel when adv is NULL
    if adv != CU_NULL
        V === nothing && error("GAE needs a critic")
        v_s = value(V, view(𝒟.cols[:s], :, 1:n))
        v_sp = CUDA.zeros(Float32, 1, n)
        chk(ccall(sym(:crux_value_next), Int32, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int64, CuPtr{Float32}),
                  V.h, 𝒟.cols[:sp], 𝒟.cols[:s], v_s, T, N, v_sp), ctx().h)
    end
    chk(ccall(sym(:crux_fill_gae_returns), Int32,
              (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{UInt8}, CuPtr{UInt8}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int64, Float32, Float32, CuPtr{Float32}, CuPtr{Float32}),
              ctx().h, 𝒟.cols[:r], 𝒟.cols[:done], 𝒟.cols[:episode_end], v_s, v_sp, T, N, γ, isnan(λ) ? 0f0 : λ, adv, ret), ctx().h)
end
"`whiten(v)` utils.jl:41-42 in place (Bessel std, no ϵ); the statistics are all-reduced when several ranks train together"
whiten!(x::Union{CuArray{Float32},SubArray{Float32}}) = chk(ccall(sym(:crux_whiten), Int32, (Ptr{Cvoid}, CuPtr{Float32}, Int64), ctx().h, x, length(x)), ctx().h)
