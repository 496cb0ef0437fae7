./tools/micro/dmma_dfma
