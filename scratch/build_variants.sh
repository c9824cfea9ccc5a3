#!/bin/bash
# A/B builds of libpmstep.so that differ in compile-time knobs of pm_particles.cu; run from the repo root.
# usage: scratch/build_variants.sh name "-DFOO=1 -DBAR=2" [name2 "..."] ...
set -e
SRC=cosmological_particle_mesh_simulation_b200/csrc
OUT=scratch/variants
mkdir -p $OUT
make -C $SRC -j4 > /dev/null
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Iinclude -I$SRC \
     -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $defs --fmad=false -c $SRC/pm_particles.cu -o $OUT/pm_particles_$name.o 2> $OUT/pm_particles_$name.ptxas.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libpmstep_$name.so \
     $SRC/pm_api.o $OUT/pm_particles_$name.o $SRC/pm_sort.o $SRC/pm_poisson.o $SRC/pm_fft.o $SRC/pm_slab.o $SRC/pm_migrate.o $SRC/pm_ic.o \
     -L/usr/local/cuda/lib64 -lcufft -Xlinker -rpath,/usr/local/cuda/lib64
  rm -f $OUT/pm_particles_$name.o
  echo "built $OUT/libpmstep_$name.so ($defs)"
done
