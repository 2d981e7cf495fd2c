# ---- solve(::OffPolicySolver, envs) (src/model_free/off_policy.jl:113-150), value_training (:66-111) for DQN and SAC ----------------------------------------
up(k, n) = cld(k, n) * n            # vector envs advance n transitions per vector step

"one DQN `value_training` call: ΔN epochs of rand! -> dqn_target -> (priorities) -> td_loss train!, then the target update (off_policy.jl:66-111, rl/dqn.jl:4-6)"
function value_training_dqn(𝒮::OffPolicySolver, q::DevDiscrete, q⁻::DevDiscrete, 𝒟::DevBuffer, buffer::DevBuffer, γ::Float32, count::Ref{Int}; seed::Integer=0)
    c = 𝒮.c_opt
    B, nA = 𝒟.capacity, length(q.outputs)
    infos = []
    for epoch in 1:c.epochs
        count[] += 1
        rand!(𝒟, buffer; i=𝒮.i, seed=seed, ctr=2 * count[])
        y = CUDA.zeros(Float32, 1, B)
        qsp = value(q⁻.q, 𝒟.cols[:sp])
        chk(ccall(sym(:crux_dqn_target), Int32, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{UInt8}, CuPtr{Float32}, Int64, Int32, Float32, CuPtr{Float32}),
                  ctx().h, 𝒟.cols[:r], 𝒟.cols[:done], qsp, B, nA, γ, y), ctx().h)
        if buffer.prioritized                                                       # off_policy.jl:83
            qs = value(q.q, 𝒟.cols[:s]); qsa = CUDA.zeros(Float32, 1, B); td = CUDA.zeros(Float32, 1, B)
            chk(ccall(sym(:crux_discrete_q_sa), Int32, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Float32}), ctx().h, qs, 𝒟.cols[:a], B, nA, qsa), ctx().h)
            chk(ccall(sym(:crux_td_error), Int32, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}), ctx().h, qsa, y, B, td), ctx().h)
            update_priorities!(buffer, indices_dev(𝒟), vec(td))
        end
        if (epoch - 1) % c.update_every == 0                                        # off_policy.jl:91-93
            info = zeros(Float32, 3)
            w = (buffer.prioritized && haskey(𝒟, :weight)) ? pointer(𝒟.cols[:weight]) : CU_NULL
            chk(ccall(sym(:crux_dqn_train), Int32, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Ptr{Float32}),
                      q.q.h, 𝒟.cols[:s], 𝒟.cols[:a], y, w, B, info), ctx().h)
            push!(infos, Dict{Any,Any}(string(c.name, "loss") => info[1], string(c.name, "grad_norm") => info[2], "Qavg" => info[3]))
        end
    end
    polyak_average!(q⁻.q, q.q, 0.005f0)                                            # off_policy.jl:108 with the default target_update (:55)
    Crux.aggregate_info(infos)
end

"one SAC `value_training` call: ΔN epochs of rand! -> target -> temperature -> double-Q critic -> actor -> polyak, each ONE library call (rl/sac.jl:4-52)"
function value_training_sac(𝒮::OffPolicySolver, st::Ptr{Cvoid}, 𝒟::DevBuffer, buffer::DevBuffer, γ::Float32, count::Ref{Int}; seed::Integer=0)
    infos = []
    for epoch in 1:𝒮.c_opt.epochs
        count[] += 1
        rand!(𝒟, buffer; i=𝒮.i, seed=seed, ctr=2 * count[])
        info = zeros(Float32, 8)
        chk(ccall(sym(:crux_sac_train), Int32,
                  (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{UInt8}, Int64, Float32, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
                   UInt64, UInt64, CuPtr{Float32}, Ptr{Float32}),
                  st, 𝒟.cols[:s], 𝒟.cols[:a], 𝒟.cols[:sp], 𝒟.cols[:r], 𝒟.cols[:done], 𝒟.capacity, γ, CU_NULL, CU_NULL, CU_NULL, seed ⊻ 0xC3C33C3C00020002, 3 * count[],
                   CU_NULL, info), ctx().h)
        push!(infos, Dict{Any,Any}("temp_loss" => info[1], "critic_loss" => info[2], "critic_grad_norm" => info[3], "actor_loss" => info[4],
                                  "actor_grad_norm" => info[5], :entropy => info[6], "Q1avg" => info[7], "Q2avg" => info[8]))
    end
    Crux.aggregate_info(infos)
end

"""
    solve(𝒮::OffPolicySolver, envs::Vector{<:MDP})

`POMDPs.solve(𝒮::OffPolicySolver, mdp)` (off_policy.jl:113-150) over a vector of env streams for the two off-policy solvers of the
hot path: DQN (`π::DiscreteNetwork`, ϵ-greedy `π_explore`) and SAC (`π::ActorCritic{SquashedGaussianPolicy, DoubleNetwork}`).
The replay buffer (`𝒮.buffer_size` rows, prioritized if `𝒮.buffer` is) lives on the device.
"""
function solve(𝒮::OffPolicySolver, envs::Vector{<:MDP})
    π, N = 𝒮.agent.π, length(envs)
    γ = Float32(discount(envs[1]))
    prioritized = Crux.isprioritized(𝒮.buffer)
    pp = 𝒮.buffer.priority_params
    buffer = DevBuffer(𝒮.S, 𝒮.agent.space, Crux.capacity(𝒮.buffer), Symbol.(Crux.extra_columns(𝒮.buffer)); prioritized,
                       α=prioritized ? pp.α : 0.6f0, β=prioritized ? pp.β : (i) -> 0.5f0)
    𝒟 = buffer_like(buffer, capacity=𝒮.c_opt.batch_size)                            # off_policy.jl:115
    s = VecSampler(envs, 𝒮.S; max_steps=𝒮.max_steps)
    count = Ref(0)
    istart = 𝒮.i
    if π isa DiscreteNetwork
        q, q⁻ = mirror(π), mirror(𝒮.agent.π⁻)
        set_adam!(q.q, 𝒮.c_opt.optimizer)
        ϵ = 𝒮.agent.π_explore isa MixedPolicy ? 𝒮.agent.π_explore.ϵ : error("CruxB200: DQN needs an ϵ-greedy π_explore (MixedPolicy)")
        collect! = (n, i) -> steps!(s, q, ϵ, buffer; Nsteps=n, i=i)      # anonymous: named local methods must not be defined per branch
        train! = () -> value_training_dqn(𝒮, q, q⁻, 𝒟, buffer, γ, count)
        back = () -> (pull!(q); pull!(q⁻); nothing)
        return offpolicy_loop(𝒮, N, istart, collect!, train!, back)
    elseif π isa ActorCritic && π.A isa SquashedGaussianPolicy && π.C isa DoubleNetwork
        ss = sac_session(𝒮)
        ne = 𝒮.agent.π_explore isa GaussianNoiseExplorationPolicy ? 𝒮.agent.π_explore : error("CruxB200: SAC needs a GaussianNoiseExplorationPolicy π_explore (rl/sac.jl:81)")
        prioritized && error("CruxB200: prioritized replay is wired for DQN only (off_policy.jl:83)")
        collect! = (n, i) -> steps!(s, ss.g, ne, buffer; Nsteps=n, i=i)
        train! = () -> value_training_sac(𝒮, ss.st, 𝒟, buffer, γ, count)
        back = () -> (pull!(ss.g); pull!(ss.c); pull_log_alpha!(𝒮, ss.st); nothing)
        try
            return offpolicy_loop(𝒮, N, istart, collect!, train!, back)
        finally
            ccall(sym(:crux_sac_destroy), Int32, (Ptr{Cvoid},), ss.st)
        end
    end
    error("CruxB200: solve(::OffPolicySolver, envs) supports DQN and SAC policies (got $(typeof(π)))")
end

"SAC: device twins of actor, critics and target critics + the fused update state (`crux_sac_create`)"
function sac_session(𝒮::OffPolicySolver)
    π, π⁻ = 𝒮.agent.π, 𝒮.agent.π⁻
    g = mirror(π.A); c = mirror(π.C); c⁻ = mirror(π⁻.C)
    set_adam!(g.mu, 𝒮.a_opt.optimizer); set_adam!(c.N1, 𝒮.c_opt.optimizer); set_adam!(c.N2, 𝒮.c_opt.optimizer)
    temp_opt = first(values(𝒮.param_optimizers)).optimizer                          # Adam of sac_temp_loss (rl/sac.jl:97)
    st = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall(sym(:crux_sac_create), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float32, Float32, Float64, Float32, Ref{Ptr{Cvoid}}),
              g.h, c.N1.h, c.N2.h, c⁻.N1.h, c⁻.N2.h, Float32(𝒮.𝒫[:SAC_log_α][1]), Float32(𝒮.𝒫[:SAC_H_target]), temp_opt.eta, 0.005f0, st), ctx().h)
    (g=g, c=c, c⁻=c⁻, st=st[])
end
"copy log α back into 𝒫[:SAC_log_α] (the reference trains it in place)"
function pull_log_alpha!(𝒮::OffPolicySolver, st::Ptr{Cvoid})
    v = Ref{Float32}(0)
    chk(ccall(sym(:crux_sac_log_alpha), Int32, (Ptr{Cvoid}, Ref{Float32}), st, v), ctx().h)
    𝒮.𝒫[:SAC_log_α][1] = v[]
end

"the interaction loop shared by the off-policy solvers (off_policy.jl:119-149); vector envs round every sample count up to whole vector steps"
function offpolicy_loop(𝒮::OffPolicySolver, N::Int, istart::Int, collect!, train!, back)
    Nfill = max(0, 𝒮.buffer_init - 0)                                               # the device buffer starts empty
    if Nfill > 0
        𝒮.i += up(Nfill, N)
        collect!(up(Nfill, N), 𝒮.i)                                                 # the initial fill counts toward N (off_policy.jl:122-133)
    end
    𝒮.log !== nothing && (back(); log(𝒮.log, 𝒮.i, Dict(), 𝒮=𝒮))
    ΔN = up(𝒮.ΔN, N)
    for 𝒮.i in range(𝒮.i, stop=istart + 𝒮.N - ΔN, step=ΔN)
        info = Dict()
        collect!(ΔN, 𝒮.i)
        𝒮.pre_train_callback(𝒮, info=info)
        training_info = train!()
        if 𝒮.log !== nothing
            Crux.elapsed(𝒮.i + 1:𝒮.i + ΔN, 𝒮.log.period) && back()
            log(𝒮.log, 𝒮.i + 1:𝒮.i + ΔN, training_info, info, 𝒮=𝒮)
        end
    end
    𝒮.i += ΔN
    back()
    check_flags()
    𝒮.agent.π
end
