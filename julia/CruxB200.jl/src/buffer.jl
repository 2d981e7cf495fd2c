# ---- ExperienceBuffer on the device (src/experience_buffer.jl) -------------------------------------------------------------------------------------
# Column ids are small integers chosen here; the order is fixed so that two buffers of one run agree on them.
const COLUMN_IDS = Dict{Symbol,Int32}(k => Int32(i - 1) for (i, k) in enumerate(
    [:s, :a, :sp, :r, :done, :episode_end, :return, :logprob, :xlogprob, :advantage, :cost, :cost_advantage, :cost_return, :value, :weight,
     :importance_weight, :t, :i, :id, :fail, :expert, :s0, :x, :var_prob, :cvar_prob, :f]))

"column schema of `mdp_data` (experience_buffer.jl:4-35): element type, row length, fill value; pixel observations stay UInt8"
function schema(S::Crux.AbstractSpace, A::Crux.AbstractSpace, extras::Vector{Symbol})
    sT = Crux.type(S) == UInt8 ? UInt8 : Float32
    sd, ad = prod(Crux.dim(S)), prod(Crux.dim(A))
    sch = Pair{Symbol,Tuple{DataType,Int,Float64}}[:s => (sT, sd, 0.0), :a => (Float32, ad, 0.0), :sp => (sT, sd, 0.0), :r => (Float32, 1, 0.0),
                                                   :done => (UInt8, 1, 0.0), :episode_end => (UInt8, 1, 0.0)]
    for k in extras
        if k in (:return, :logprob, :xlogprob, :advantage, :cost, :cost_advantage, :cost_return, :value, :var_prob, :cvar_prob, :f)
            push!(sch, k => (Float32, 1, 0.0))
        elseif k in (:weight, :importance_weight)
            push!(sch, k => (Float32, 1, 1.0))
        elseif k in (:fail, :expert)
            push!(sch, k => (UInt8, 1, 0.0))
        elseif k in (:t, :i, :id)
            push!(sch, k => (Int64, 1, 0.0))
        elseif k == :s0
            push!(sch, k => (sT, sd, 0.0))
        elseif k == :x
            push!(sch, k => (Float32, ad, 0.0))
        else
            error("CruxB200: unrecognized column $k")
        end
    end
    sch
end
dtype_code(::Type{UInt8}) = CRUX_U8
dtype_code(::Type{Float32}) = CRUX_F32
dtype_code(::Type{Int32}) = CRUX_I32
dtype_code(::Type{Int64}) = CRUX_I64

mutable struct DevBuffer
    h::Ptr{Cvoid}
    sch::Vector{Pair{Symbol,Tuple{DataType,Int,Float64}}}
    cols::Dict{Symbol,CuArray}          # whole columns `[rowlen, capacity]`, zero-copy views of the library's allocations
    capacity::Int
    prioritized::Bool
    α::Float32
    β::Function
    indices::Vector{Int}                # 1-based ids of the last sample (filled lazily from the device copy)
end
"`ExperienceBuffer(S, A, capacity, extras; prioritized, priority_params...)` experience_buffer.jl:74-80"
function DevBuffer(S::Crux.AbstractSpace, A::Crux.AbstractSpace, capacity::Int, extras::Vector{Symbol}=Symbol[]; prioritized::Bool=false,
                   α::Real=0.6f0, β::Function=(i) -> 0.5f0)
    prioritized && !(:weight in extras) && (extras = vcat(extras, :weight))
    DevBuffer(schema(S, A, extras), capacity; prioritized, α, β)
end
function DevBuffer(sch, capacity::Int; prioritized::Bool=false, α::Real=0.6f0, β::Function=(i) -> 0.5f0)
    descs = [ColDesc(COLUMN_IDS[k], dtype_code(T), n, init) for (k, (T, n, init)) in sch]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall(sym(:crux_buffer_create), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{ColDesc}, Int32, Float32, Ref{Ptr{Cvoid}}),
              ctx().h, capacity, length(descs), descs, prioritized ? 1 : 0, α, out), ctx().h)
    cols = Dict{Symbol,CuArray}()
    for (k, (T, n, _)) in sch
        p, rl, dt = Ref{CuPtr{Cvoid}}(CU_NULL), Ref{Int64}(0), Ref{Int32}(0)
        chk(ccall(sym(:crux_buffer_col), Int32, (Ptr{Cvoid}, Int32, Ref{CuPtr{Cvoid}}, Ref{Int64}, Ref{Int32}), out[], COLUMN_IDS[k], p, rl, dt), ctx().h)
        cols[k] = unsafe_wrap(CuArray, reinterpret(CuPtr{T}, p[]), (n, capacity))
    end
    b = DevBuffer(out[], collect(sch), cols, capacity, prioritized, Float32(α), β, Int[])
    finalizer(b -> (b.h != C_NULL && ccall(sym(:crux_buffer_destroy), Int32, (Ptr{Cvoid},), b.h); b.h = C_NULL), b)
end
"`buffer_like(b; capacity)` experience_buffer.jl:82-85"
Crux.buffer_like(b::DevBuffer; capacity::Int=b.capacity, kwargs...) = DevBuffer(b.sch, capacity; prioritized=b.prioritized, α=b.α, β=b.β)

function state(b::DevBuffer)
    e, n, t, c = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    chk(ccall(sym(:crux_buffer_state), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), b.h, e, n, t, c), ctx().h)
    (elements=Int(e[]), next_ind=Int(n[]) + 1, total_count=Int(t[]), capacity=Int(c[]))     # next_ind 1-based like the reference
end
Base.length(b::DevBuffer) = state(b).elements
Crux.capacity(b::DevBuffer) = b.capacity
Crux.isprioritized(b::DevBuffer) = b.prioritized
Base.keys(b::DevBuffer) = first.(b.sch)
Base.haskey(b::DevBuffer, k::Symbol) = haskey(b.cols, k)
Crux.extra_columns(b::DevBuffer) = [k for k in keys(b) if !(k in (:s, :a, :sp, :r, :done, :episode_end))]
"`b[:key]`: a VIEW of the first `elements` rows (experience_buffer.jl:173); callers write through it (`𝒟[:advantage] .= ...`, ppo.jl:61)"
Base.getindex(b::DevBuffer, k::Symbol) = view(b.cols[k], :, 1:length(b))
Crux.clear!(b::DevBuffer) = (chk(ccall(sym(:crux_buffer_clear), Int32, (Ptr{Cvoid},), b.h), ctx().h); b)

"`push!(b, data; ids)` experience_buffer.jl:232-259: `data` a Dict of host Arrays or of CuArrays (`[rowlen, N]`); returns the ring indices"
function Base.push!(b::DevBuffer, data::AbstractDict{Symbol}; ids=nothing)
    ks = [k for k in keys(b) if haskey(data, k)]
    isempty(ks) && return Int[]
    N = ids === nothing ? size(data[ks[1]], ndims(data[ks[1]])) : length(ids)
    on_host = !(data[ks[1]] isa CuArray)
    colids = Int32[COLUMN_IDS[k] for k in ks]
    conv(k, x) = (T = first(Dict(b.sch)[k]); eltype(x) == T ? x : T.(x))
    arrays = [conv(k, data[k]) for k in ks]                                     # kept alive until the call returns
    first_ind = Ref{Int64}(0)
    GC.@preserve arrays begin
        ptrs = Ptr{Cvoid}[on_host ? Ptr{Cvoid}(pointer(x)) : reinterpret(Ptr{Cvoid}, pointer(x)) for x in arrays]
        ids0 = ids === nothing ? nothing : (on_host ? Int32.(ids .- 1) : CuArray(Int32.(ids .- 1)))
        idp = ids0 === nothing ? C_NULL : (on_host ? Ptr{Cvoid}(pointer(ids0)) : reinterpret(Ptr{Cvoid}, pointer(ids0)))
        GC.@preserve ids0 chk(ccall(sym(:crux_buffer_push), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}, Ref{Int64}),
                                    b.h, N, length(ks), colids, ptrs, on_host ? 1 : 0, idp, first_ind), ctx().h)
    end
    mod1.(first_ind[] + 1:first_ind[] + N, b.capacity)
end
"rows `next_ind : next_ind+N-1` were written in place through the column views (zero-copy rollouts): advance the ring like `push!`"
commit_rows!(b::DevBuffer, N::Int) =
    chk(ccall(sym(:crux_buffer_push), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}, Ptr{Int64}), b.h, N, 0, C_NULL, C_NULL, 0, C_NULL, C_NULL), ctx().h)

"`get_last_N_indices(b, N)` experience_buffer.jl:223-229"
function Crux.get_last_N_indices(b::DevBuffer, N::Int)
    out = Vector{Int64}(undef, max(1, min(N, b.capacity))); n = Ref{Int64}(0)
    chk(ccall(sym(:crux_buffer_last_n_indices), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ref{Int64}), b.h, min(N, b.capacity), out, n), ctx().h)
    Int.(out[1:n[]]) .+ 1
end

"`split_batches(N, fracs)` experience_buffer.jl:126-131"
function split_batches(N::Int, fracs)
    out = Vector{Int64}(undef, length(fracs))
    chk(ccall(sym(:crux_split_batches), Int32, (Int64, Ptr{Float64}, Int32, Ptr{Int64}), N, Float64.(collect(fracs)), length(fracs), out))
    Int.(out)
end
"`rand!(target, sources...; i, fracs)` experience_buffer.jl:303-315 (uniform or prioritized per source; `target.indices` = the LAST source's ids)"
function Random.rand!(target::DevBuffer, sources::DevBuffer...; i=1, fracs=ones(length(sources)) ./ length(sources), seed::Integer=0, ctr::Integer=0)
    clear!(target)
    Bs = split_batches(target.capacity, fracs)
    for (k, (src, B)) in enumerate(zip(sources, Bs))
        if src.prioritized
            chk(ccall(sym(:crux_buffer_sample_prioritized), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Float32, Int32, Ptr{Float64}, UInt64, UInt64),
                      target.h, src.h, B, Float32(src.β(i)), COLUMN_IDS[:weight], C_NULL, seed, ctr + 2 * (k - 1)), ctx().h)
        else
            chk(ccall(sym(:crux_buffer_sample_uniform), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Int32}, UInt64, UInt64),
                      target.h, src.h, B, C_NULL, seed, ctr + 2 * (k - 1)), ctx().h)
        end
    end
    target
end
"device int32 ids (0-based) of the last sample: what `update_priorities!` consumes without a host round trip"
function indices_dev(b::DevBuffer)
    p, n = Ref{CuPtr{Int32}}(CU_NULL), Ref{Int64}(0)
    chk(ccall(sym(:crux_buffer_indices), Int32, (Ptr{Cvoid}, Ref{CuPtr{Int32}}, Ref{Int64}), b.h, p, n), ctx().h)
    unsafe_wrap(CuArray, p[], Int(n[]))
end
"`update_priorities!(b, I, v)` experience_buffer.jl:290-301 (`I` the device ids of `indices_dev`, `v = |Q(s,a) - y|` on the device)"
Crux.update_priorities!(b::DevBuffer, I::CuArray{Int32}, v::CuArray{Float32}) =
    chk(ccall(sym(:crux_buffer_update_priorities), Int32, (Ptr{Cvoid}, CuPtr{Int32}, CuPtr{Float32}, Int64), b.h, I, v, length(I)), ctx().h)
"`episodes(b)` experience_buffer.jl:194-221 from the `episode_end` flags"
function Crux.episodes(b::DevBuffer)
    ee = vec(Array(b[:episode_end])) .!= 0
    stops = findall(ee)
    starts = vcat(1, stops[1:end-1] .+ 1)
    collect(zip(starts, stops))
end
