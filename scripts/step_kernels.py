"""One eager training step of the paired-tower harness between cudaProfilerStart/Stop, for an ncu launch list:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file step.csv python scripts/step_kernels.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pmt_learning_for_semantic_segmentation_and_disparity_b200 import harness, sharding

world = sharding.World(0, 0, 1, None)
step, model = harness.build_training_step(world, batch_per_gpu=4, sync_bn=True, cuda_graph=False, paired_tower=True)
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
