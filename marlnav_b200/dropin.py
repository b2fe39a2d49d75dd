"""Run the reference's own CLI with the fused CUDA ``Env`` dropped in.

    python -m marlnav_b200.dropin -rc -sn 0          # == python -m marlnav -rc -sn 0
    python -m marlnav_b200.dropin -np 1024 -nt 1024000

Needs the reference package ``marlnav`` importable (it is NOT part of this repository).  The only
thing changed is the name ``marlnav.environment.Env`` -- ``marlnav/__main__.py:7`` imports it from
there -- so argument parsing, MAPPO, reward-check and rendering are the reference's own code.
"""
import runpy
import sys


def main():
    try:
        import marlnav.environment as ref_env
    except ImportError as e:       # pragma: no cover - depends on the user's installation
        raise SystemExit(f"the reference package `marlnav` is not importable: {e}")
    import marlnav_b200
    ref_env.Env = marlnav_b200.Env
    runpy.run_module("marlnav", run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    sys.exit(main())
