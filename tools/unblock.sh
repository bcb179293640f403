#!/usr/bin/env bash
# Unblock-day triage of a staged reference tree: see tools/unblock.py.  Usage: tools/unblock.sh [TREE]
exec python "$(dirname "$0")/unblock.py" "$@"
