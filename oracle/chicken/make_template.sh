#!/bin/bash
# Regenerates oracle/chicken/template_config.json with the reference's own preprocessor (build container only:
# needs /root/reference).  The template is what chkn-prep writes for oracle/chicken/box.py; make_job.py scales it.
set -e
here=$(cd "$(dirname "$0")" && pwd)
tmp=$(mktemp -d)
cp "$here/box.py" "$tmp/"
(cd "$tmp" && PYTHONPATH=/root/reference/src/lib python /root/reference/src/chicken/chkn_prep.py --job=box --binary)
cp "$tmp/box/config.json" "$here/template_config.json"
rm -rf "$tmp"
