#!/bin/bash
timeout 300 python scripts/ingest_bench.py --subjects 3 2>&1 | tail -1 | cut -c1-300
timeout 300 python scripts/train_all_subjects.py --mat-dir /tmp/eav_mat --subjects 3 --epochs 2 2>&1 | tail -1 | cut -c1-500
timeout 300 python scripts/train_all_subjects.py --mat-dir /tmp/eav_mat --subjects 3 --epochs 2 --legacy-order 2>&1 | tail -1 | cut -c1-500
timeout 300 python scripts/train_all_subjects.py --subjects 6 --epochs 2 --separable --lr 1e-3 2>&1 | tail -1 | cut -c1-500
