#!/bin/bash
# build_variant.sh NAME "-DFOO=1 -DBAR=2": builds gnn_ecommerce_b200/variants/liblgc_NAME.so with extra nvcc flags
# (select it at run time with LGC_B200_LIB=...). The default library is rebuilt afterwards by build().
set -e
NAME=$1; shift
mkdir -p gnn_ecommerce_b200/variants
LGC_NVCC_EXTRA="$*" python -m gnn_ecommerce_b200.build --force > /dev/null
cp gnn_ecommerce_b200/liblgc_b200.so gnn_ecommerce_b200/variants/liblgc_$NAME.so
echo built $NAME
