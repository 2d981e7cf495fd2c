"""CPU oracle package: TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
The product package ``crux.jl_b200`` never imports this."""
