# ---- library, status codes, structs that cross the ABI by value / by reference (include/crux_cuda.h) ----------------------------------
const lib = Ref{Ptr{Cvoid}}(C_NULL)
function __init__()
    path = get(ENV, "LIBCRUX_CUDA", joinpath(@__DIR__, "..", "..", "..", "crux.jl_b200", "lib", "libcrux_cuda.so"))
    lib[] = Libdl.dlopen(path)                       # throws when the CUDA library is missing: there is no CPU fallback
    v = ccall(sym(:crux_abi_version), Int32, ())
    v == 1 || error("libcrux_cuda.so speaks ABI version $v, this package binds version 1")
end
sym(s::Symbol) = Libdl.dlsym(lib[], s)

struct CruxError <: Exception
    code::Int32
    msg::String
end
Base.showerror(io::IO, e::CruxError) = print(io, "libcrux_cuda status ", e.code, ": ", e.msg)

const CRUX_ERR_NAN = Int32(3)
"status -> exception.  CRUX_ERR_NAN is the reference's `error(\"NaN detected! Loss: ...\")` (training.jl:20)."
function chk(rc::Int32, ctx::Ptr{Cvoid}=C_NULL)
    rc == 0 && return nothing
    msg = unsafe_string(ccall(sym(:crux_last_error), Cstring, (Ptr{Cvoid},), ctx))
    rc == CRUX_ERR_NAN ? error("NaN detected! $msg") : throw(CruxError(rc, msg))
end

# C layouts.  sizeof / field offsets are asserted in test/runtests.jl against the numbers `tests/abi_smoke layout` prints.
struct PPOHp                        # crux_ppo_hp
    eps_clip::Float32; lambda_p::Float32; lambda_e::Float32; target_kl::Float32
    a2c::Int32; actor_epochs::Int32; actor_batch::Int32; critic_epochs::Int32; critic_batch::Int32
    actor_max_batches::Int64; critic_max_batches::Int64
end
struct LagrangeHp                   # crux_lagrange_hp
    target_cost::Float32; penalty_max::Float32; Ki_max::Float32; Ki::Float32; Kp::Float32; Kd::Float32
    ema_alpha::Float64
    cost_epochs::Int32
    cost_batch::Int64; cost_max_batches::Int64
end
struct RolloutCols                  # crux_rollout_cols (device pointers)
    s::CuPtr{Float32}; a::CuPtr{Float32}; sp::CuPtr{Float32}; r::CuPtr{Float32}
    done::CuPtr{UInt8}; episode_end::CuPtr{UInt8}
    logprob::CuPtr{Float32}
end
struct ColDesc                      # crux_col_desc
    id::Int32; dtype::Int32; rowlen::Int64; init::Float64
end
const CRUX_U8, CRUX_F32, CRUX_I32, CRUX_I64 = Int32(0), Int32(1), Int32(2), Int32(3)
const INFO_LOSS, INFO_GRAD_NORM, INFO_ENTROPY, INFO_KL, INFO_CLIP, INFO_AVG_ADV, INFO_AVG_RET, INFO_VALID = 1:8   # 1-based rows of an info record

# ---- context: replaces device / gpucall / cpucall / mdcall (src/devices.jl:1-21): nothing hops between devices per call ----------------
mutable struct Ctx
    h::Ptr{Cvoid}
end
function Ctx(dev::Integer=CUDA.deviceid(CUDA.device()))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    # share CUDA.jl's task-local stream so that CuArray operations and library launches are ordered without extra synchronisation
    chk(ccall(sym(:crux_ctx_create), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), dev, reinterpret(Ptr{Cvoid}, CUDA.stream().handle), out))
    c = Ctx(out[])
    finalizer(c -> (c.h != C_NULL && ccall(sym(:crux_ctx_destroy), Int32, (Ptr{Cvoid},), c.h); c.h = C_NULL), c)
end
const CTX = Ref{Union{Nothing,Ctx}}(nothing)
ctx() = (CTX[] === nothing && (CTX[] = Ctx()); CTX[]::Ctx)
sync(c::Ctx=ctx()) = chk(ccall(sym(:crux_ctx_sync), Int32, (Ptr{Cvoid},), c.h), c.h)
"sticky device-side NaN flag (gradients / advantages): throws like `training.jl:20` / `sampler.jl:270`"
check_flags(c::Ctx=ctx()) = chk(ccall(sym(:crux_ctx_check), Int32, (Ptr{Cvoid},), c.h), c.h)

"pinned host array (cudaMallocHost) wrapped as a Julia Array: the env callbacks write into it, the device reads it over PCIe"
function pinned(::Type{T}, dims::Int...) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall(sym(:crux_pinned_alloc), Int32, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx().h, prod(dims) * sizeof(T), p), ctx().h)
    unsafe_wrap(Array, Ptr{T}(p[]), dims; own=false)
end
