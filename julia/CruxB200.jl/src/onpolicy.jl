# ---- solve(::OnPolicySolver, envs) (src/model_free/on_policy.jl:80-109) with policy_gradient_training (:56-78) as ONE library call ---------------------
"""
    target_kl(es)

The KL threshold of a `TrainingParams.early_stopping` closure.  PPO / A2C / REINFORCE build it as `(infos) -> infos[end][:kl] > target_kl`
(rl/ppo.jl:59, a2c.jl:46, reinforce.jl:32); the fused update evaluates exactly that rule on the device after every minibatch, so the
threshold is recovered from the closure itself by bisection over Float32 (`Inf32` when the closure never stops, e.g. the default
`(info) -> false`, training.jl:8).  A closure that is not a monotone threshold on `:kl` cannot be run on the device: error.
"""
function target_kl(es)
    probe(x) = try Bool(es([Dict{Any,Any}(:kl => x)])) catch; try Bool(es(Dict{Any,Any}(:kl => x))) catch; error("CruxB200: early_stopping must be a threshold on info[:kl]") end end
    probe(Inf32) || return Inf32
    probe(-Inf32) && error("CruxB200: early_stopping stops unconditionally")
    lo, hi = -floatmax(Float32), floatmax(Float32)      # invariant: !probe(lo), probe(hi)
    probe(0f0) ? (hi = 0f0) : (lo = 0f0)
    while nextfloat(lo) < hi
        mid = lo / 2 + hi / 2
        (mid <= lo || mid >= hi) && break
        probe(mid) ? (hi = mid) : (lo = mid)
    end
    lo                                                  # kl > lo stops: lo is the largest KL that still continues
end

actor_mlp(g::DevGaussian) = g.mu
actor_mlp(g::DevCategorical) = g.d.q
loss_kind(f) = f === Crux.ppo_loss ? :ppo : f === Crux.a2c_loss ? :a2c : f === Crux.reinforce_loss ? :reinforce :
               error("CruxB200: the fused on-policy update implements ppo_loss, a2c_loss and reinforce_loss (got $f)")
maxb(p) = isinf(p.max_batches) ? Int64(0) : Int64(p.max_batches)

"`policy_gradient_training(𝒮, 𝒟)` on_policy.jl:56-78: batch_train!(actor) then batch_train!(critic) (training.jl:28-55) -> info Dict"
function policy_gradient_training(𝒮::OnPolicySolver, g::Union{DevGaussian,DevCategorical}, V::Union{DevMLP,Nothing}, 𝒟::DevBuffer; orders=(nothing, nothing), seed::Integer=rand(UInt64))
    isempty(𝒮.param_optimizers) || error("CruxB200: param_optimizers are not supported by the fused on-policy update")
    𝒮.cost_opt === nothing || error("CruxB200: cost critics (LagrangePPO) go through crux_lagrange_ppo_update; bind it the same way")
    n = length(𝒟)
    a, c, 𝒫 = 𝒮.a_opt, 𝒮.c_opt, 𝒮.𝒫
    kind = loss_kind(a.loss)
    weight = kind == :reinforce ? 𝒟.cols[:return] : 𝒟.cols[:advantage]          # reinforce_loss = a2c head with the return as the weight
    hp = Ref(PPOHp(get(𝒫, :ϵ, 0.2f0), kind == :reinforce ? 1f0 : get(𝒫, :λp, 1f0), kind == :reinforce ? 0f0 : get(𝒫, :λe, 0.1f0), target_kl(a.early_stopping),
                   kind == :ppo ? 0 : 1, a.epochs, a.batch_size, c === nothing ? 0 : c.epochs, c === nothing ? 1 : c.batch_size, maxb(a), c === nothing ? 0 : maxb(c)))
    nmb_a, nmb_c = cld(n, a.batch_size), c === nothing ? 1 : cld(n, c.batch_size)
    ia = zeros(Float32, 8, max(1, a.epochs * nmb_a)); ic = zeros(Float32, 8, max(1, (c === nothing ? 0 : c.epochs) * nmb_c))
    ord(o) = o === nothing ? CU_NULL : pointer(o)
    GC.@preserve orders chk(ccall(sym(:crux_ppo_update), Int32,
              (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Ref{PPOHp},
               CuPtr{Int32}, CuPtr{Int32}, UInt64, Ptr{Float32}, Ptr{Float32}),
              g.h, V === nothing ? C_NULL : V.h, 𝒟.cols[:s], 𝒟.cols[:a], 𝒟.cols[:logprob], weight, 𝒟.cols[:return], n, hp,
              ord(orders[1]), ord(orders[2]), seed, ia, ic), ctx().h)
    # batch_train! pushes the SAME Dict for every minibatch (training.jl:43): the reference's aggregates are the last trained minibatch's values
    info = Dict{Any,Any}()
    la = findlast(>(0), ia[INFO_VALID, :])
    if la !== nothing
        info[string(a.name, "loss")] = ia[INFO_LOSS, la]; info[string(a.name, "grad_norm")] = ia[INFO_GRAD_NORM, la]
        info[:entropy] = ia[INFO_ENTROPY, la]; info[:kl] = ia[INFO_KL, la]
        kind == :ppo && (info[:clip_fraction] = ia[INFO_CLIP, la]; info[:avg_advantage] = ia[INFO_AVG_ADV, la]; info[:avg_return] = ia[INFO_AVG_RET, la])
    end
    lc = c === nothing ? nothing : findlast(>(0), ic[INFO_VALID, :])
    lc !== nothing && (info[string(c.name, "loss")] = ic[INFO_LOSS, lc]; info[string(c.name, "grad_norm")] = ic[INFO_GRAD_NORM, lc])
    info
end

"""
    solve(𝒮::OnPolicySolver, envs::Vector{<:MDP})

The loop of `POMDPs.solve(𝒮::OnPolicySolver, mdp)` (on_policy.jl:80-109) over `length(envs)` independent env streams: rollouts, GAE,
callbacks and the update run on the device; the trained parameters are copied back into `𝒮.agent.π` before returning (and before every
log call, so `LoggerParams.fns` evaluate the current policy with stock Crux code).
"""
function solve(𝒮::OnPolicySolver, envs::Vector{<:MDP})
    π = 𝒮.agent.π
    A = π isa ActorCritic ? π.A : π
    (A isa GaussianPolicy || A isa DiscreteNetwork) ||
        error("CruxB200: the on-policy path supports GaussianPolicy(μ, logΣ vector) and DiscreteNetwork (categorical) actors (got $(typeof(A)))")
    g = A isa DiscreteNetwork ? DevCategorical(A) : mirror(A)
    V = (π isa ActorCritic && 𝒮.c_opt !== nothing) ? mirror(π.C) : nothing
    set_adam!(actor_mlp(g), 𝒮.a_opt.optimizer)
    V !== nothing && set_adam!(V, 𝒮.c_opt.optimizer)
    N = length(envs)
    𝒮.ΔN % N == 0 || error("ΔN = $(𝒮.ΔN) must be a multiple of the $N env streams")
    T = 𝒮.ΔN ÷ N
    𝒟 = DevBuffer(𝒮.S, 𝒮.agent.space, 𝒮.ΔN, Symbol.(𝒮.required_columns))          # on_policy.jl:82
    γ, λ = Float32(discount(envs[1])), 𝒮.λ_gae
    s = VecSampler(envs, 𝒮.S; max_steps=𝒮.max_steps, λ=λ)
    sync_back() = (pull!(g); V !== nothing && pull!(V); nothing)
    𝒮.log !== nothing && isnothing(𝒮.log.sampler) && (𝒮.log.sampler = Sampler(envs[1], 𝒮.agent, S=𝒮.S, required_columns=𝒮.required_columns, λ=λ, max_steps=𝒮.max_steps))
    𝒮.log !== nothing && log(𝒮.log, 𝒮.i, 𝒮=𝒮)                                      # pre-train performance (on_policy.jl:88)
    for 𝒮.i in range(𝒮.i, stop=𝒮.i + 𝒮.N - 𝒮.ΔN, step=𝒮.ΔN)
        info = Dict()
        clear!(𝒟)
        steps!(s, g, 𝒟; Nsteps=𝒮.ΔN, reset=true)                                   # on_policy.jl:96
        fill_gae_returns!(𝒟, V, T, N, γ, λ)                                        # terminate_episode! for all closed ranges at once
        𝒮.post_sample_callback(𝒟, info=info, 𝒮=𝒮)
        𝒮.interaction_storage !== nothing && push!(𝒮.interaction_storage, Dict(k => Array(𝒟[k]) for k in keys(𝒟)))
        𝒮.post_batch_callback(𝒟, info=info, 𝒮=𝒮)                                   # PPO: 𝒟[:advantage] .= whiten(𝒟[:advantage]) on CuArray views
        training_info = policy_gradient_training(𝒮, g, V, 𝒟)
        if 𝒮.log !== nothing
            Crux.elapsed(𝒮.i + 1:𝒮.i + 𝒮.ΔN, 𝒮.log.period) && sync_back()
            log(𝒮.log, 𝒮.i + 1:𝒮.i + 𝒮.ΔN, training_info, info, 𝒮=𝒮)
        end
    end
    𝒮.i += 𝒮.ΔN
    sync_back()
    check_flags()
    𝒮.agent.π
end

"`whiten` of a device view (PPO's post_batch_callback broadcasts `whiten(𝒟[:advantage])`): one fused device pass, Bessel std as utils.jl:41-42"
function Crux.whiten(v::SubArray{Float32,2,<:CuArray})
    out = copy(v)
    whiten!(out)
    out
end
