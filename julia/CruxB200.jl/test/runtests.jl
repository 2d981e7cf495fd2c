# Run with:  LIBCRUX_CUDA=/path/to/crux.jl_b200/lib/libcrux_cuda.so julia --project -e 'using Pkg; Pkg.test()'
# (Julia is not part of the build image of this repository: these tests have been written against include/crux_cuda.h and the numbers
#  `tests/_build/abi_smoke layout` prints; tests/test_abi_symbols.py checks that the sizes below still match the C structs.)
using Test, CruxB200, Crux, Flux, POMDPs, POMDPModels, CUDA

@testset "C struct layouts (tests/abi_smoke.c layout)" begin
    @test sizeof(CruxB200.PPOHp) == 56
    @test fieldoffset(CruxB200.PPOHp, 5) == 16 && fieldoffset(CruxB200.PPOHp, 10) == 40 && fieldoffset(CruxB200.PPOHp, 11) == 48
    @test sizeof(CruxB200.LagrangeHp) == 56
    @test fieldoffset(CruxB200.LagrangeHp, 7) == 24 && fieldoffset(CruxB200.LagrangeHp, 8) == 32 && fieldoffset(CruxB200.LagrangeHp, 9) == 40
    @test sizeof(CruxB200.RolloutCols) == 56
    @test sizeof(CruxB200.ColDesc) == 24
    @test fieldoffset(CruxB200.ColDesc, 3) == 8 && fieldoffset(CruxB200.ColDesc, 4) == 16
end

@testset "target_kl recovered from the early_stopping closure" begin
    @test CruxB200.target_kl((infos) -> infos[end][:kl] > 0.012f0) == 0.012f0
    @test CruxB200.target_kl((infos) -> infos[end][:kl] > 0.015) == prevfloat(Float32(0.015)) || CruxB200.target_kl((infos) -> infos[end][:kl] > 0.015) == Float32(0.015)
    @test CruxB200.target_kl((info) -> false) == Inf32
end

if CUDA.functional()
    mlp(i, o) = Chain(Dense(i, 64, tanh), Dense(64, 64, tanh), Dense(64, o))
    @testset "network mirror: value / exploration / logpdf agree with Flux on the CPU" begin
        π = ActorCritic(GaussianPolicy(ContinuousNetwork(mlp(17, 6)), -0.5f0 * ones(Float32, 6)), ContinuousNetwork(mlp(17, 1)))
        g, V = mirror(π.A), mirror(π.C)
        s = randn(Float32, 17, 257)
        @test Array(value(V, CuArray(s))) ≈ value(π.C, s) rtol = 1e-5 atol = 1e-5
        ϵ = randn(Float32, 6, 257)
        a, lp = exploration(g, CuArray(s); eps=CuArray(ϵ))
        @test Array(a) ≈ ϵ .* exp.(-0.5f0) .+ value(π.A.μ, s) rtol = 1e-5 atol = 1e-5
        @test Array(lp) ≈ logpdf(π.A, s, Array(a)) rtol = 1e-5 atol = 3e-5
        @test entropy(g, CuArray(s)) ≈ entropy(π.A, s)
    end
    @testset "PPO on 64 LinQuad streams: runs, trains, parameters come back" begin
        π = ActorCritic(GaussianPolicy(ContinuousNetwork(mlp(17, 6)), -0.5f0 * ones(Float32, 6)), ContinuousNetwork(mlp(17, 1)))
        before = deepcopy(Flux.params(π.A.μ.network)[1])
        𝒮 = PPO(π=π, S=ContinuousSpace(17), N=3 * 512, ΔN=512, max_steps=50, a_opt=(epochs=2, batch_size=128), c_opt=(epochs=2, batch_size=128),
                log=(period=512, fns=[], verbose=false))
        m = LinQuadMDP()
        @test solve(𝒮, [m for _ in 1:64]) === π
        @test 𝒮.i == 3 * 512
        @test Flux.params(π.A.μ.network)[1] != before
        env = DeviceLinQuad(m, 64; max_steps=50)
        @test solve(𝒮, env) === π
    end
    @testset "PPO with a categorical actor on 16 grid-world streams (examples/rl/cartpole.jl:8-9,17-25)" begin
        mdp = SimpleGridWorld(size=(10, 10), tprob=0.7)
        S = state_space(mdp)
        as = [actions(mdp)...]
        A = DiscreteNetwork(Chain(Dense(Crux.dim(S)..., 64, relu), Dense(64, 64, relu), Dense(64, length(as))), as)
        V = ContinuousNetwork(Chain(Dense(Crux.dim(S)..., 64, relu), Dense(64, 64, relu), Dense(64, 1)))
        before = deepcopy(Flux.params(A.network)[1])
        𝒮 = PPO(π=ActorCritic(A, V), S=S, N=3 * 256, ΔN=256, max_steps=30, a_opt=(epochs=2, batch_size=64), c_opt=(epochs=2, batch_size=64),
                log=(period=256, fns=[], verbose=false))
        solve(𝒮, [mdp for _ in 1:16])
        @test 𝒮.i == 3 * 256
        @test Flux.params(A.network)[1] != before
        c = CruxB200.DevCategorical(A)
        s = CUDA.rand(Float32, 2, 100)
        idx, oh, lp = exploration(c, s; seed=1, ctr=0)
        @test all(sum(Array(oh), dims=1) .== 1) && all(Array(lp) .<= 0)
        @test Array(lp) ≈ logpdf(A, Array(s), Array(oh)) rtol = 1e-5 atol = 1e-5
    end
    @testset "replay buffer ring semantics (test/experience_buffer_tests.jl:121-147)" begin
        b = DevBuffer(ContinuousSpace(2), ContinuousSpace(1), 5)
        d(n, v) = Dict{Symbol,Any}(:s => fill(Float32(v), 2, n), :a => fill(Float32(v), 1, n), :sp => fill(Float32(v), 2, n), :r => fill(Float32(v), 1, n),
                                   :done => zeros(UInt8, 1, n), :episode_end => zeros(UInt8, 1, n))
        @test push!(b, d(3, 1)) == [1, 2, 3]
        @test push!(b, d(3, 2)) == [4, 5, 1]
        @test length(b) == 5 && CruxB200.state(b).next_ind == 2
        @test vec(Array(b[:r])) == Float32[2, 1, 1, 2, 2]
        @test CruxB200.split_batches(100, [1 / 3, 1 / 3, 1 / 3]) == [34, 33, 33]
    end
end
