# Builds libitcpd_b200 for sm_100a -- the nvcc analogue of the reference's deps/build.jl:15-17
# (which runs `gcc -O3 -fPIC -shared` on the two sparse-sign C files).
# Usage from the package root:  julia deps/build_b200.jl   (nvcc must be on PATH; no GPU needed to build)
csrc = joinpath(@__DIR__, "..", "..", "csrc")
out  = joinpath(@__DIR__, "..", "..", "lib")
mkpath(out)
srcs = [joinpath(csrc, f) for f in sort(filter(endswith(".cu"), readdir(csrc)))]   # every translation unit (same set as csrc/build.sh)
lib_file = joinpath(out, "libitcpd_b200.so")
compile = `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared
           -o $lib_file $srcs -cudart static -ldl -lpthread -lrt`
run(compile)
