"""Command-line options with the reference's flag names (parsers/opts.py:15-44), on stdlib argparse
(configargparse is not available).  ``--config FILE`` accepts the same "key = value" lines.
Additions for this back end: ``--arch b200`` (default), ``--gpus``, ``--seed``, ``--max_bounce``
(the reference can only take max_bounce from the XML), ``--save_hdr``, ``--spp_per_launch``."""
import argparse
import shlex

__all__ = ["get_options"]


def _expand_config(argv):
    """configargparse-style: '--config path' injects 'key = value' lines as '--key value' defaults."""
    if argv is None:
        import sys
        argv = sys.argv[1:]
    argv = list(argv)
    if "--config" not in argv:
        return argv
    i = argv.index("--config")
    path = argv[i + 1]
    injected = []
    with open(path, "r") as fh:
        for line in fh:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            if "=" in line:
                k, v = line.split("=", 1)
            else:
                k, v = (line.split(None, 1) + [""])[:2]
            k, v = k.strip(), v.strip()
            if v.lower() in ("true", ""):
                injected.append(f"--{k}")
            elif v.lower() == "false":
                continue
            else:
                injected.extend([f"--{k}"] + shlex.split(v))
    # command line wins over the file
    return injected + argv[:i] + argv[i + 2:]


def get_options(delayed_parse: bool = False, argv=None):
    parser = argparse.ArgumentParser(description="B200-native drop-in for AdaPT's `--type pt` / `--type vpt` renderers")
    parser.add_argument("--config", help="Config file path", type=str, default=None)
    parser.add_argument("--iter_num", default=-1, help="Number of iterations (-1 means from XML / 2000)", type=int)
    parser.add_argument("--normalize", default=0.0, help="Normalize the output picture with its <x> quantile value", type=float)
    parser.add_argument("--output_freq", default=0, help="Whether to output intermediate results (0 means no)", type=int)
    parser.add_argument("--input_path", default="./scenes/", help="Input scene file folder", type=str)
    parser.add_argument("--output_path", default="./outputs/", help="Output image file folder", type=str)
    parser.add_argument("--chkpt_path", default="./checkpoint/", help="Check-point file folder", type=str)
    parser.add_argument("--img_name", default="pbr", help="Output image name", type=str)
    parser.add_argument("--img_ext", default="png", choices=["png", "jpg", "bmp"], help="Output image extension", type=str)
    parser.add_argument("--scene", default="cbox", help="Name of the scene", type=str)
    parser.add_argument("--name", default="complex.xml", help="Scene file name with extension", type=str)
    parser.add_argument("--arch", default="b200", choices=["b200", "cuda", "gpu"], help="Backend (always the sm_100a CUDA path)")
    parser.add_argument("--save_iter", default=-1, type=int, help="Iteration to save check-point")
    parser.add_argument("--type", default="vpt", choices=["vpt", "pt", "bdpt", "ao"],
                        help="Algorithm to be used (defaults as in the reference, parsers/opts.py:28-31; `pt` and `vpt` run on the device, the others are refused)")
    parser.add_argument("-p", "--profile", default=False, action="store_true", help="Print per-stage device timings")
    parser.add_argument("--no_gui", default=False, action="store_true", help="Accepted for compatibility (there is no GUI)")
    parser.add_argument("-d", "--debug", default=False, action="store_true", help="Accepted for compatibility")
    parser.add_argument("-a", "--analyze", default=False, action="store_true", help="Accepted for compatibility")
    parser.add_argument("-l", "--load", default=False, action="store_true", help="Whether to load checkpoint")
    parser.add_argument("--no_cache", default=False, action="store_true", help="Accepted for compatibility")
    parser.add_argument("--no_save_fig", default=False, action="store_true", help="Whether to save images")
    parser.add_argument("--no_watermark", default=False, action="store_true", help="Whether to add watermark")
    # additions
    parser.add_argument("--gpus", default=1, type=int, help="Number of GPUs (launch through torchrun for > 1)")
    parser.add_argument("--seed", default=0, type=int, help="Seed of the counter-based RNG")
    parser.add_argument("--max_bounce", default=-1, type=int, help="Override the XML max_bounce")
    parser.add_argument("--save_hdr", default=False, action="store_true", help="Also dump the HDR mean buffer as .npy")
    parser.add_argument("--spp_per_launch", default=0, type=int, help="spp batched per device launch (0 = auto)")
    if delayed_parse:
        return parser
    return parser.parse_args(_expand_config(argv))
