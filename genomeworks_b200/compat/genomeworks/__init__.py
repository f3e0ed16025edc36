"""`genomeworks` import alias for callers written against pygenomeworks: put genomeworks_b200/compat on sys.path and
`import genomeworks.cudapoa`, `genomeworks.cudaaligner`, `genomeworks.cuda` resolve to the B200 engine's mirrors
(pygenomeworks/genomeworks/{cuda,cudapoa,cudaaligner}). Only the GPU-backed modules of the hot path are aliased; the reference's
pure-Python utilities (simulators, io, utilities) are out of scope."""
