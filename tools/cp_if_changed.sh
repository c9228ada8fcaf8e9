#!/bin/sh
# cp_if_changed.sh <src> <dst>: copy only when the contents differ (keeps dst's time stamp otherwise)
cmp -s "$1" "$2" || cp "$1" "$2"
