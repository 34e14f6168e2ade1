# Builds libitcpd_b200 for sm_100a -- the nvcc analogue of the reference's deps/build.jl:15-17
# (which runs `gcc -O3 -fPIC -shared` on the two sparse-sign C files).
# Usage from the package root:  julia deps/build_b200.jl   (nvcc must be on PATH; no GPU needed to build)
include(joinpath(@__DIR__, "b200_paths.jl"))
println("built ", B200Paths.build())
