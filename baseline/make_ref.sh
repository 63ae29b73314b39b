#!/usr/bin/env bash
# Recreates baseline/_ref/: a verbatim copy of the reference files the two training steps import.
# baseline/_ref/ is git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored,
# so it travels to the GPU box, where /root/reference does not exist. Build container only.
set -euo pipefail
SRC=${XVA_REFERENCE_ROOT:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
rm -rf "$DST"
mkdir -p "$DST/python/fastpitch1_1/fastpitch" "$DST/python/fastpitch1_1/common" "$DST/python/hifigan"
for f in model transformer attention alignment loss_function attn_loss_function; do
  cp "$SRC/python/fastpitch1_1/fastpitch/$f.py" "$DST/python/fastpitch1_1/fastpitch/"
done
for f in layers stft audio_processing utils; do
  cp "$SRC/python/fastpitch1_1/common/$f.py" "$DST/python/fastpitch1_1/common/"
done
cp "$SRC/python/fastpitch1_1/lamb.py" "$DST/python/fastpitch1_1/"
for f in models meldataset utils env; do
  cp "$SRC/python/hifigan/$f.py" "$DST/python/hifigan/"
done
cp "$SRC/python/hifigan/config_v1.json" "$DST/python/hifigan/"
# xVAPitch --hifi_only step (posterior encoder, waveform decoder, discriminator, losses) and what those modules import
mkdir -p "$DST/python/xvapitch"
for f in model hifigan wavenet losses audio util glow_tts sdp; do
  cp "$SRC/python/xvapitch/$f.py" "$DST/python/xvapitch/"
done
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/REFERENCE_COMMIT"
echo "baseline/_ref: $(find "$DST" -type f | wc -l) files"
