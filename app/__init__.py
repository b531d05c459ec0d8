"""`app` import path of the reference (harveybc/gym-fx: app/env.py, app/plugin_loader.py, app/config.py), served by
the B200-native implementation in gym_fx_b200, so that `from app.env import GymFxEnv` (tools/smoke_test.py:27,
app/main.py:10 of the reference) keeps working.  Only the modules of the env.step() path exist here; the CLI,
config-file merging and remote-config glue of the reference's `app/` are out of scope."""
