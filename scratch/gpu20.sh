#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scratch/prof_slab.py > gpurun_out/t_prof.log 2>&1; echo rc=$?; tail -12 gpurun_out/t_prof.log
