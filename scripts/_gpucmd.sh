python scripts/gemm_ab.py deformcontact_b200/libdcb200_r01gemm.so deformcontact_b200/libdcb200.so 2>&1 | tail -6
for cfg in "0 0" "1 0" "0 2.6e-8" "1 2.6e-8" "1 3.5e-8"; do
  set -- $cfg
  echo "=== LO_RN=$1 COMP=$2 (DRAIN_KB=4)"
  DCB200_T2_LO_RN=$1 DCB200_T2_COMP=$2 python scripts/gemm_lab.py 2>&1 | grep -E "layer fwd|layer dW|scores|attn @" | cut -c1-120
  DCB200_T2_LO_RN=$1 DCB200_T2_COMP=$2 python scripts/acc_lab.py 2>&1 | tail -1
  LAB_GROUP=2 DCB200_T2_LO_RN=$1 DCB200_T2_COMP=$2 python scripts/acc_lab.py 2>&1 | tail -1
done
echo "=== DRAIN_KB=8 LO_RN=1 COMP=2.6e-8"
DCB200_DRAIN_KB=8 DCB200_DRAIN_KB_SHORT=8 python scripts/gemm_lab.py 2>&1 | grep -E "layer fwd|scores|attn @" | cut -c1-120
DCB200_DRAIN_KB=8 DCB200_DRAIN_KB_SHORT=8 python scripts/acc_lab.py 2>&1 | tail -1
LAB_GROUP=2 DCB200_DRAIN_KB=8 DCB200_DRAIN_KB_SHORT=8 python scripts/acc_lab.py 2>&1 | tail -1
