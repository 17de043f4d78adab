"""Times BASELINE config 3 (gudepthcompnet18 + cbr3-c1 + frozen GED loss net, B=16, 240x320) and prints the per-kind kernel time table
of one profiled step.  Usage: python tools/time_gud.py [batch]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
print(json.dumps(bench.gudepthcompnet_training(dev, B=B, graph=False)))
print(json.dumps(bench.gudepthcompnet_training(dev, B=B, graph=True)))
