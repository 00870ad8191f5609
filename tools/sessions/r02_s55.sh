#!/bin/bash
# Round-2 session 55: full ncu capture of the one-group k_logic in its shipped shape (128-thread blocks, eight per SM), one lane
ADAPT_LANES=1 bash tools/ncu_any.sh k_logic logic128 --also ''
ls -la gpurun_out/prof_logic128.ncu-rep
