# Launch-bounds scan of the general element kernel (compiled on the GPU box): edits the MAXT/MINB columns of one line of
# general_dispatch.hpp, rebuilds, times the configuration.  usage: bash tools/gpu_bounds_scan.sh
set -u
D=mrhyde_b200/csrc/general_dispatch.hpp
cp $D /tmp/dispatch.orig
scan() {  # $1 = sed pattern prefix (unique), $2 = bench args, $3 = kernel name pattern in the ptxas log, rest = "MAXT,MINB" variants
  local pat="$1" args="$2" kpat="$3"; shift 3
  for v in "$@"; do
    cp /tmp/dispatch.orig $D
    local maxt=${v%,*} minb=${v#*,}
    sed -i "s/\(${pat}\)[0-9]*, [0-9]*)/\1${maxt}, ${minb})/" $D
    grep -n "${pat}" $D | head -1
    make -s -j8 -C mrhyde_b200/csrc 2>&1 | grep -E "error" | head -3
    python tools/ptxas_report.py mrhyde_b200/csrc/build/general.ptxas.log | grep "false" | grep -E "$kpat" | head -1
    eval timeout 300 python tools/bench_general.py $args 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   bounds ${v}: %.3f ms  %.1f M elem/s' % (d['ms_per_assemble'], d['elements_per_s']/1e6))"
  done
}
scan 'X("navier stokes", 3, 1, 8, 4, 1, GenNs31, ' "ns 64" "NavierStokesPhys<3" "128,2" "128,4" "256,1"
scan 'X("linearelasticity", 3, 1, 8, 4, 1, GenLe31, ' "le 64" "ElasticityPhys<3, 1>" "256,2" "128,4"
scan 'X("maxwell", 3, 1, 8, 4, 1, MaxwellPhys, ' "maxwell 48" "MaxwellPhys" "128,2" "128,4"
cp /tmp/dispatch.orig $D
