#!/bin/bash
timeout 300 python tools/latency_graph.py
