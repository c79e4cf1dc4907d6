"""Drop-in for the reference's match_single_ui.py (the script the GUI runs over SSH,
GUI/depth_estimation_connect_server_GUI.py:256): same CLI (-g, -i, -f) as match_single.py, but the pair is read from
./UI_use/ (match_single_ui.py:30) and the map is written as uint8 * 2 (:55)."""
from __future__ import annotations

from . import match_single


def main(argv=None):
    argv = list(argv) if argv is not None else None
    import sys

    args = sys.argv[1:] if argv is None else argv
    extra = []
    if not any(a.startswith("--image-dir") for a in args):
        extra += ["--image-dir", "./UI_use/"]
    if not any(a.startswith("--scale") for a in args):
        extra += ["--scale", "2"]
    match_single.main(list(args) + extra)


if __name__ == "__main__":
    main()
