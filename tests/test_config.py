"""CPU tests of the yaml loader against the reference algorithm (hcpdiff/utils/utils.py:43-72).

omegaconf is not installable offline, so the expected values below are the result of running the reference's
`load_config_with_cli` BY HAND on the files this test writes:

    def load_config(path, remove_undefined=True):
        cfg = OmegaConf.load(path)
        if '_base_' in cfg:
            for base in cfg['_base_']:
                cfg = OmegaConf.merge(load_config(base, remove_undefined=False), cfg)     # <- `cfg` (file + EARLIER bases) wins
            del cfg['_base_']
        ...
    load_config_with_cli: cfg = merge(load_config(path, False), from_cli(args)); then remove every key whose value is '---'
"""
import os
import textwrap

from hcp_diffusion_b200.utils.config import load_config_with_cli


def _write(d, name, body):
    p = os.path.join(d, name)
    with open(p, "w") as f:
        f.write(textwrap.dedent(body))
    return p


def test_three_bases_earlier_base_wins_and_undefined_sentinel(tmp_path):
    d = str(tmp_path)
    _write(d, "grand.yaml", """
        train: {lr: 1.0, steps: 10, from_grand: true}
        model: {name: grand}
    """)
    _write(d, "a.yaml", f"""
        _base_: [{d}/grand.yaml]
        train: {{lr: 2.0, a_only: 1}}
        logger: {{kind: a, every: 5}}
        drop_me: {{x: 1}}
    """)
    _write(d, "b.yaml", """
        train: {lr: 3.0, b_only: 2, steps: 30}
        logger: {kind: b}
        data: [1, 2, 3]
        drop_me: '---'
    """)
    _write(d, "c.yaml", """
        train: {lr: 4.0, c_only: 3, steps: 40, a_only: 99}
        data: [9]
        extra: {deep: {v: 1, w: '---'}}
    """)
    top = _write(d, "top.yaml", f"""
        _base_: [{d}/a.yaml, {d}/b.yaml, {d}/c.yaml]
        train: {{steps: 50, resume: '---'}}
        extra: {{deep: {{u: 7}}}}
    """)
    cfg = load_config_with_cli(top, ["train.cli=5", "logger.every=---", "model.name=fromcli"])
    # top wins over every base; among bases the EARLIER one wins (a > b > c); a's own base (grand) loses to a
    assert cfg.train.steps == 50                      # file
    assert cfg.train.lr == 2.0                        # a (first base) beats b, c and grand
    assert cfg.train.a_only == 1                      # a beats c's 99
    assert cfg.train.b_only == 2 and cfg.train.c_only == 3 and cfg.train.from_grand is True
    assert cfg.logger.kind == "a"
    assert cfg.data == [1, 2, 3]                      # lists are replaced, not merged: b (earlier) beats c
    assert cfg.model.name == "fromcli" and cfg.train.cli == 5
    # '---': a defined drop_me first, so a's dict wins over b's sentinel (b is merged UNDER the accumulated cfg)
    assert cfg.drop_me == {"x": 1}
    # the sentinel set in the file / on the command line survives every merge and is removed at the very end
    assert "resume" not in cfg.train
    assert "every" not in cfg.logger
    assert cfg.extra.deep == {"u": 7, "v": 1}         # c's w: '---' is removed, u from the file, v from c
    assert "_base_" not in cfg


def test_sentinel_in_file_deletes_inherited_key(tmp_path):
    d = str(tmp_path)
    _write(d, "base.yaml", """
        model: {ema: {decay: 0.99}, wd: 0.1}
    """)
    top = _write(d, "top.yaml", f"""
        _base_: [{d}/base.yaml]
        model: {{ema: '---'}}
    """)
    cfg = load_config_with_cli(top, [])
    assert "ema" not in cfg.model and cfg.model.wd == 0.1


def test_relative_base_and_resolvers(tmp_path):
    d = str(tmp_path)
    _write(d, "base.yaml", """
        a: {b: 3}
        dtype: ${hcp.dtype:bf16}
    """)
    top = _write(d, "top.yaml", """
        _base_: [base.yaml]
        c: ${a.b}
        e: ${hcp.eval:"2*3"}
    """)
    cfg = load_config_with_cli(top, ["a.b=4"])
    assert cfg.c == 4 and cfg.e == 6 and cfg.dtype == "torch.bfloat16"
