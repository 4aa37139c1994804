#!/bin/bash
# build the extension from anywhere; non-zero exit (and the compiler output) when it fails
set -o pipefail
cd /root/repo && python multimodal_clinical_b200/build.py "$@" 2>&1 | grep -E "error|Error|warning: v|_lf_fusion.so|failed" | tail -20
