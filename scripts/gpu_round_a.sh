#!/bin/bash
bash scripts/gpu_lookup_check.sh
