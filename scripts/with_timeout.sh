#!/bin/bash
# with_timeout.sh SECONDS cmd...: run cmd in its own session and kill the whole session on timeout
T=$1; shift
setsid "$@" &
pid=$!
( sleep "$T"; kill -KILL -- -"$pid" 2>/dev/null ) &
w=$!
wait "$pid"; rc=$?
kill "$w" 2>/dev/null
exit $rc
