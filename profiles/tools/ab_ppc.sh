#!/bin/bash
cd "$(dirname "$0")/../.."
for so in gpurun_ab/*.so; do echo $so; PPCSEQ_B200_LIB=$PWD/$so python profiles/tools/ppc_rates.py; done
