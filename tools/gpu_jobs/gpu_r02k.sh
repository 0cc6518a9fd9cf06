#!/bin/bash
TAG=${1:-r02k}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; env "$@" timeout 200 python tools/debug_nan_poison.py $EXTRA 2>&1 | grep -E "replay|Error|error" | head -8; }
EXTRA="" run graphs_default X=1
EXTRA="nographs" run eager_default X=1
EXTRA="" run no_early_zero RLIPV2_EARLY_ZERO=0
EXTRA="" run no_short_attn RLIPV2_SHORT_ATTN=0
EXTRA="" run no_lang_stream RLIPV2_LANG_STREAM=0
EXTRA="" run no_value_stream RLIPV2_VALUE_STREAM=0
EXTRA="" run no_pos_stream RLIPV2_POS_STREAM=0
EXTRA="" run no_wgrad_stream RLIPV2_WGRAD_STREAM=0
EXTRA="" run no_flag_wait RLIPV2_FLAG_WAIT=0
EXTRA="" run no_fused_prologue RLIPV2_MSDA_FUSED_PROLOGUE=0
EXTRA="" run no_small_ops RLIPV2_SMALL_OPS=0
EXTRA="" run no_text_stream RLIPV2_TEXT_STREAM=0
EXTRA="" run no_parallel_losses RLIPV2_PARALLEL_LOSSES=0
EXTRA="" run no_fused_box RLIPV2_FUSED_BOX_LOSS=0
EXTRA="" run no_stacked_heads RLIPV2_STACKED_HEADS=0
EXTRA="" run no_fused_conv RLIPV2_FUSED_CONV=0
