import sys, os
sys.path.insert(0, ".")
import torch
from scripts.dev_time_layer import time_layer
