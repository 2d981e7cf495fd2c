# ---- networks: a Flux Chain of Dense layers mirrored on the device (src/policies.jl:68-98) -----------------------------------------------
actcode(f) = (f === tanh || f === Flux.NNlib.tanh_fast) ? Int32(1) : f === Flux.relu ? Int32(2) : f === identity ? Int32(0) :
             error("CruxB200: unsupported activation $f (tanh, relu and identity are fused)")
dense_layers(c::Flux.Chain) = [l for l in c.layers if l isa Flux.Dense]
"`Flux.params` order, every `W` as `vec` of the Julia `[out, in]` matrix followed by `b`"
flat(c::Flux.Chain) = reduce(vcat, [vcat(vec(Float32.(Flux.cpu(l.weight))), Float32.(Flux.cpu(l.bias))) for l in dense_layers(c)])

mutable struct DevMLP
    h::Ptr{Cvoid}
    dims::Vector{Int32}
    chain::Flux.Chain        # the user's Flux model: `pull!` writes the trained parameters back into it
end
function DevMLP(c::Flux.Chain)
    ls = dense_layers(c)
    length(ls) == length(c.layers) || error("CruxB200: only Chain(Dense...) networks are mirrored (got $(typeof.(c.layers)))")
    dims = Int32[size(ls[1].weight, 2); [size(l.weight, 1) for l in ls]]
    acts = Int32[actcode(l.σ) for l in ls]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall(sym(:crux_mlp_create), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}), ctx().h, length(acts), dims, acts, out), ctx().h)
    m = DevMLP(out[], dims, c)
    push_params!(m)
    finalizer(m -> (m.h != C_NULL && ccall(sym(:crux_mlp_destroy), Int32, (Ptr{Cvoid},), m.h); m.h = C_NULL), m)
end
push_params!(m::DevMLP) = chk(ccall(sym(:crux_mlp_set_params), Int32, (Ptr{Cvoid}, Ptr{Float32}), m.h, flat(m.chain)), ctx().h)
function num_params(m::DevMLP)
    n = Ref{Int64}(0)
    chk(ccall(sym(:crux_mlp_num_params), Int32, (Ptr{Cvoid}, Ref{Int64}), m.h, n), ctx().h)
    n[]
end
"copy the trained parameters back into the user's Flux model (BSON.@save, plotting, `value(π, s)` on the CPU keep working)"
function pull!(m::DevMLP)
    v = Vector{Float32}(undef, num_params(m))
    chk(ccall(sym(:crux_mlp_get_params), Int32, (Ptr{Cvoid}, Ptr{Float32}), m.h, v), ctx().h)
    o = 0
    for l in dense_layers(m.chain)
        nw = length(l.weight); copyto!(l.weight, reshape(v[o+1:o+nw], size(l.weight))); o += nw
        nb = length(l.bias);   copyto!(l.bias, v[o+1:o+nb]); o += nb
    end
    m.chain
end
"`Adam(η, β, ϵ)` of a TrainingParams (training.jl:3): Flux's Float64 scalars, moments reset"
function set_adam!(m::DevMLP, opt)
    opt isa Flux.Optimise.Adam || error("CruxB200: the fused update implements Flux.Adam (got $(typeof(opt)))")
    chk(ccall(sym(:crux_mlp_set_adam), Int32, (Ptr{Cvoid}, Float64, Float64, Float64, Float64), m.h, opt.eta, opt.beta[1], opt.beta[2], opt.epsilon), ctx().h)
end

"`value(π::ContinuousNetwork, s)` policies.jl:94 for a `[features, B]` CuArray"
function Crux.value(m::DevMLP, s::CuArray{Float32})
    B = size(s, ndims(s))
    y = CUDA.zeros(Float32, Int(m.dims[end]), B)
    chk(ccall(sym(:crux_mlp_forward), Int32, (Ptr{Cvoid}, CuPtr{Float32}, Int64, CuPtr{Float32}), m.h, s, B, y), ctx().h)
    y
end
"`value(π, s, a) = network(vcat(s, a))` policies.jl:96"
function Crux.value(m::DevMLP, s::CuArray{Float32}, a::CuArray{Float32})
    B = size(s, ndims(s))
    y = CUDA.zeros(Float32, Int(m.dims[end]), B)
    chk(ccall(sym(:crux_mlp_forward_sa), Int32, (Ptr{Cvoid}, CuPtr{Float32}, Int32, CuPtr{Float32}, Int32, Int64, CuPtr{Float32}),
              m.h, s, size(s, 1), a, size(a, 1), B, y), ctx().h)
    y
end
"`polyak_average!(to, from, τ)` policies.jl:48-59 and `copyto!(to, from)` :61-65"
Crux.polyak_average!(to::DevMLP, from::DevMLP, τ=1f0) = chk(ccall(sym(:crux_mlp_polyak), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Float32), to.h, from.h, τ), ctx().h)
Base.copyto!(to::DevMLP, from::DevMLP) = (chk(ccall(sym(:crux_mlp_copy), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), to.h, from.h), ctx().h); to)

# ---- GaussianPolicy(μ, logΣ) policies.jl:315-350, SquashedGaussianPolicy :355-400 -----------------------------------------------------------
mutable struct DevGaussian
    h::Ptr{Cvoid}
    mu::DevMLP
    adim::Int
    policy                   # the user's policy (its logΣ vector is written back by pull!)
end
const_vector(n::ContinuousNetwork) = (l = n.network.layers[1]; l isa ConstantLayer ? Float32.(vec(Flux.cpu(l.vec))) : nothing)
function DevGaussian(π::Union{GaussianPolicy,SquashedGaussianPolicy})
    mu = DevMLP(π.μ.network)
    ls = const_vector(π.logΣ)
    squashed = π isa SquashedGaussianPolicy
    if ls === nothing     # [μ | logΣ] heads on one trunk are passed as ONE Chain with 2A outputs (half_cheetah_mujoco.jl:37-43)
        error("CruxB200: a state-dependent logΣ network must share μ's trunk: build the policy with one Chain of 2A outputs and pass it as μ")
    end
    out = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall(sym(:crux_gaussian_create), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float32}, Int32, Float32, Ref{Ptr{Cvoid}}),
              ctx().h, mu.h, length(ls), ls, squashed ? 1 : 0, squashed ? π.ascale : 1f0, out), ctx().h)
    g = DevGaussian(out[], mu, length(ls), π)
    finalizer(g -> (g.h != C_NULL && ccall(sym(:crux_gaussian_destroy), Int32, (Ptr{Cvoid},), g.h); g.h = C_NULL), g)
end
function pull!(g::DevGaussian)
    pull!(g.mu)
    p = Ref{CuPtr{Float32}}(CU_NULL)
    chk(ccall(sym(:crux_gaussian_log_sigma_ptr), Int32, (Ptr{Cvoid}, Ref{CuPtr{Float32}}), g.h, p), ctx().h)
    ls = Array(unsafe_wrap(CuArray, p[], g.adim))
    l = g.policy.logΣ.network.layers[1]
    copyto!(l.vec, reshape(ls, size(l.vec)))
    g.policy
end
"`exploration(π, s)` -> `(a, logprob)` policies.jl:338-344, :388-394.  `eps` injects the standard-normal draws (parity runs)."
function Crux.exploration(g::DevGaussian, s::CuArray{Float32}; eps=nothing, seed::Integer=0, ctr::Integer=0, kwargs...)
    B = size(s, ndims(s))
    a = CUDA.zeros(Float32, g.adim, B); lp = CUDA.zeros(Float32, 1, B)
    chk(ccall(sym(:crux_gaussian_explore), Int32, (Ptr{Cvoid}, CuPtr{Float32}, Int64, CuPtr{Float32}, UInt64, UInt64, CuPtr{Float32}, CuPtr{Float32}),
              g.h, s, B, eps === nothing ? CU_NULL : eps, seed, ctr, a, lp), ctx().h)
    a, lp
end
function POMDPs.action(g::DevGaussian, s::CuArray{Float32})
    B = size(s, ndims(s)); a = CUDA.zeros(Float32, g.adim, B)
    chk(ccall(sym(:crux_gaussian_action), Int32, (Ptr{Cvoid}, CuPtr{Float32}, Int64, CuPtr{Float32}), g.h, s, B, a), ctx().h)
    a
end
function Crux.logpdf(g::DevGaussian, s::CuArray{Float32}, a::CuArray{Float32})
    B = size(s, ndims(s)); out = CUDA.zeros(Float32, 1, B)
    chk(ccall(sym(:crux_gaussian_logpdf), Int32, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}), g.h, s, a, B, out), ctx().h)
    out
end
"`entropy(π, s)`: the scalar `1.4189385332f0 + sum(logΣ)` for a GaussianPolicy (policies.jl:348), `[1, B]` for the squashed policy (:398)"
function Crux.entropy(g::DevGaussian, s::CuArray{Float32})
    B = size(s, ndims(s)); out = CUDA.zeros(Float32, 1, B)
    chk(ccall(sym(:crux_gaussian_entropy), Int32, (Ptr{Cvoid}, CuPtr{Float32}, Int64, CuPtr{Float32}), g.h, s, B, out), ctx().h)
    g.policy isa GaussianPolicy ? Array(out)[1] : out
end

# ---- DiscreteNetwork policies.jl:104-157 -----------------------------------------------------------------------------------------------------
mutable struct DevDiscrete
    q::DevMLP
    outputs::Vector
    policy::DiscreteNetwork
end
DevDiscrete(π::DiscreteNetwork) = DevDiscrete(DevMLP(π.network), collect(π.outputs), π)
pull!(d::DevDiscrete) = (pull!(d.q); d.policy)
Crux.value(d::DevDiscrete, s::CuArray{Float32}) = value(d.q, s)
"`action(π::DiscreteNetwork, s)`: argmax, the first maximum wins (policies.jl:124) -> 1-based indices on the device"
function action_index(d::DevDiscrete, s::CuArray{Float32})
    q = value(d.q, s); B = size(q, 2); idx = CUDA.zeros(Int32, B)
    chk(ccall(sym(:crux_discrete_argmax), Int32, (Ptr{Cvoid}, CuPtr{Float32}, Int64, Int32, CuPtr{Int32}), ctx().h, q, B, size(q, 1), idx), ctx().h)
    idx .+ Int32(1)
end
"ϵ-greedy `exploration(::MixedPolicy)` policies.jl:474-494 over all streams -> (1-based indices, one-hot `[nA, B]`, logprob `[1, B]`)"
function eps_greedy(d::DevDiscrete, s::CuArray{Float32}, ϵ::Real; seed::Integer=0, ctr::Integer=0)
    q = value(d.q, s); nA, B = size(q)
    idx = CUDA.zeros(Int32, B); oh = CUDA.zeros(Float32, nA, B); lp = CUDA.zeros(Float32, 1, B)
    chk(ccall(sym(:crux_discrete_eps_greedy), Int32,
              (Ptr{Cvoid}, CuPtr{Float32}, Int64, Int32, Float64, Ptr{Float64}, UInt64, UInt64, CuPtr{Int32}, CuPtr{Float32}, CuPtr{Float32}),
              ctx().h, q, B, nA, ϵ, C_NULL, seed, ctr, idx, oh, lp), ctx().h)
    idx .+ Int32(1), oh, lp
end

"""
A `DiscreteNetwork` used as an on-policy ACTOR (examples/rl/cartpole.jl:8,17-25): `ppo_loss` / `a2c_loss` / `reinforce_loss` see it through
`logpdf` = `categorical_logpdf` (policies.jl:135) and `entropy` (:152-155); `crux_categorical_create` gives the update its handle.
"""
mutable struct DevCategorical
    h::Ptr{Cvoid}
    d::DevDiscrete
end
function DevCategorical(π::DiscreteNetwork)
    d = DevDiscrete(π)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall(sym(:crux_categorical_create), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), ctx().h, d.q.h, length(d.outputs), out), ctx().h)
    c = DevCategorical(out[], d)
    finalizer(c -> (c.h != C_NULL && ccall(sym(:crux_gaussian_destroy), Int32, (Ptr{Cvoid},), c.h); c.h = C_NULL), c)
end
pull!(c::DevCategorical) = pull!(c.d)
"`exploration(π::DiscreteNetwork, s)` policies.jl:137-142 over all streams: softmax -> categorical draw -> (1-based indices, one-hot `[nA, B]`, logprob `[1, B]`)"
function Crux.exploration(c::DevCategorical, s::CuArray{Float32}; seed::Integer=0, ctr::Integer=0, kwargs...)
    q = value(c.d.q, s); nA, B = size(q)
    idx = CUDA.zeros(Int32, B); lp = CUDA.zeros(Float32, 1, B)
    chk(ccall(sym(:crux_discrete_explore), Int32,
              (Ptr{Cvoid}, CuPtr{Float32}, Int64, Int32, Ptr{Float64}, UInt64, UInt64, CuPtr{Int32}, CuPtr{Float32}),
              ctx().h, q, B, nA, C_NULL, seed, ctr, idx, lp), ctx().h)
    idx1 = idx .+ Int32(1)
    oh = Float32.(Int32.(1:nA) .== reshape(idx1, 1, B))          # Flux.onehotbatch(π, a) as a dense `[nA, B]` array on the device
    idx1, oh, lp
end

# ---- mirror(π): device twins of the policy trees the hot path supports ---------------------------------------------------------------------------
struct DevActorCritic{TA,TC}
    A::TA
    C::TC
end
struct DevDouble
    N1::DevMLP
    N2::DevMLP
end
mirror(π::ContinuousNetwork) = DevMLP(π.network)
mirror(π::Union{GaussianPolicy,SquashedGaussianPolicy}) = DevGaussian(π)
mirror(π::DiscreteNetwork) = DevDiscrete(π)
mirror(π::DoubleNetwork) = DevDouble(mirror(π.N1), mirror(π.N2))
mirror(π::ActorCritic) = DevActorCritic(mirror(π.A), mirror(π.C))
pull!(π::DevActorCritic) = (pull!(π.A); pull!(π.C); nothing)
pull!(π::DevDouble) = (pull!(π.N1); pull!(π.N2); nothing)
