# Where libitcpd_b200 lives and how it is built -- ONE definition, included by both deps/build_b200.jl and
# ext/ITCPDB200Ext/ITCPDB200Ext.jl, so that the extension looks for the library exactly where the build writes it
# (the reference keeps the same pair in deps/build.jl:1-28 and src/ITensorCPD.jl:25-36).
#
# Layout after the drop-in (INTEGRATION.md, "files to copy"), <pkg> = the ITensorCPD.jl package root:
#   <pkg>/deps/b200_paths.jl, <pkg>/deps/build_b200.jl
#   <pkg>/deps/b200/include/itcpd_b200.h
#   <pkg>/deps/b200/itensorcpd.jl_b200/csrc/*.cu, *.cuh          (sources; they include "../../include/itcpd_b200.h")
#   <pkg>/deps/b200/itensorcpd.jl_b200/lib/libitcpd_b200.so      (built here)
#   <pkg>/ext/ITCPDB200Ext/ITCPDB200Ext.jl
# In the development repository the same two files sit under itensorcpd.jl_b200/julia/{deps,ext}; there the source
# root is the repository root (three levels above this file).  ENV["ITCPD_B200_ROOT"] overrides the source root and
# ENV["ITCPD_B200_LIB"] the library file itself.
module B200Paths

const DEPS_DIR = @__DIR__

function source_root()
    haskey(ENV, "ITCPD_B200_ROOT") && return ENV["ITCPD_B200_ROOT"]
    dropin = joinpath(DEPS_DIR, "b200")
    isdir(joinpath(dropin, "itensorcpd.jl_b200", "csrc")) && return dropin
    return normpath(joinpath(DEPS_DIR, "..", "..", ".."))          # development repository
end

csrc_dir() = joinpath(source_root(), "itensorcpd.jl_b200", "csrc")
library_path() = get(ENV, "ITCPD_B200_LIB", joinpath(source_root(), "itensorcpd.jl_b200", "lib", "libitcpd_b200.so"))

# nvcc analogue of the reference's `gcc -O3 -fPIC -shared` (deps/build.jl:15-17); nvcc cross-compiles without a GPU
function build()
    csrc = csrc_dir()
    lib_file = library_path()
    mkpath(dirname(lib_file))
    srcs = [joinpath(csrc, f) for f in sort(filter(endswith(".cu"), readdir(csrc)))]   # every translation unit (same set as csrc/build.sh)
    run(`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared
         -o $lib_file $srcs -cudart static -ldl -lpthread -lrt`)
    return lib_file
end

end # module
