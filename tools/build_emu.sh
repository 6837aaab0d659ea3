#!/bin/bash
# Development aid: serial CPU emulation of the CUDA core (see loco_mujoco_b200/csrc/locosim_emu.cpp).
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch
g++ -O2 -fPIC -shared -std=c++17 -Wall -Wno-unused-function -Wno-unused-variable -Wno-unknown-pragmas -o scratch/liblocosim_emu.so loco_mujoco_b200/csrc/locosim_emu.cpp -lm
