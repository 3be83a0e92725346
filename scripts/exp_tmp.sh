nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate scripts/micro/umma_rate.cu && timeout 120 /tmp/umma_rate 2>&1 | head -26
