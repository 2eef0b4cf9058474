"""Checkpoint I/O in the reference's on-disk format (velora/utils/restore.py:107-272), so that an agent trained on the B200 kernels loads
into stock velora and vice versa (SURVEY.md sections 5 and 8f-4):

    <dir>/model_state.safetensors   "<module>.<state_dict key>"            actor / critic / critic2 / critic_target / critic2_target;
                                                                           mask buffers included, the entropy module's log_alpha is not
    <dir>/optim_state.safetensors   "<optimizer>.<parameter index>.<key>"  actor_optim / critic_optim / critic2_optim / entropy_optim
    <dir>/metadata.json             {"model": constructor keywords, "<optimizer>": param_groups, ["buffer": buffer metadata]}
    <dir>/buffer_state.safetensors  (optional) the six replay arrays

The reference re-creates the agent from metadata["model"] (an environment id -> gymnasium); here `load_state` restores INTO an agent the
caller has constructed, which is the part that matters for the hot path (parameter names, shapes and order are the contract).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict

import torch

MODULE_NAMES = ("actor", "critic", "entropy")


def _agent_state(agent) -> Dict[str, Dict[str, Any]]:
    """{"modules": {...}, "optimizers": {...}} - velora/models/base.py:296-312."""
    out: Dict[str, Dict[str, Any]] = {"modules": {}, "optimizers": {}}
    for name in MODULE_NAMES:
        module = getattr(agent, name, None)
        if module is None:
            continue
        for key, val in module.state_dict().items():
            out["optimizers" if "optim" in key else "modules"][key] = val
    return out


def save_model(agent, dirpath: str | Path, *, buffer: bool = False, force: bool = False, metadata: Dict[str, Any] | None = None) -> None:
    from safetensors.torch import save_file

    root = Path(dirpath)
    root.mkdir(parents=True, exist_ok=True)
    model_path = Path(root, "model_state").with_suffix(".safetensors")
    if not force and model_path.exists():
        raise FileExistsError(f"A model state already exists in the '{root}' directory! Either change the 'dirpath', delete the folders "
                              "contents, or use 'force=True' to allow overwriting.")
    state = _agent_state(agent)
    tensors = {f"{m}.{k}": v.detach().contiguous().cpu() for m, sd in state["modules"].items() for k, v in sd.items()}
    optim_tensors: Dict[str, torch.Tensor] = {}
    meta: Dict[str, Any] = {"model": metadata if metadata is not None else getattr(agent, "metadata", {})}
    for name, sd in state["optimizers"].items():
        for pid, pstate in sd.get("state", {}).items():
            for k, v in pstate.items():
                optim_tensors[f"{name}.{pid}.{k}"] = (v if torch.is_tensor(v) else torch.tensor(v)).detach().cpu()
        meta[name] = sd["param_groups"]
    save_file(tensors, model_path)
    save_file(optim_tensors, Path(root, "optim_state").with_suffix(".safetensors"))
    if buffer:
        meta["buffer"] = agent.buffer.metadata()
        save_file({k: v.contiguous() for k, v in agent.buffer.state_dict().items()}, Path(root, "buffer_state").with_suffix(".safetensors"))
    with Path(root, "metadata").with_suffix(".json").open("w") as f:
        f.write(json.dumps(meta, indent=2, default=str))


def load_state(agent, dirpath: str | Path, *, buffer: bool = False):
    """Restores the module and optimizer states (and optionally the buffer) saved by `save_model` - ours or the reference's - into `agent`."""
    from safetensors.torch import load_file

    root = Path(dirpath)
    paths = {k: Path(root, k).with_suffix(ext) for k, ext in (("metadata", ".json"), ("model_state", ".safetensors"), ("optim_state", ".safetensors"),
                                                                  ("buffer_state", ".safetensors"))}
    for k in ("model_state", "optim_state", "metadata"):
        if not paths[k].exists():
            raise FileNotFoundError(f"'{paths[k]}' does not exist!")
    if buffer and not paths["buffer_state"].exists():
        raise FileNotFoundError(f"Buffer state '{paths['buffer_state']}' does not exist! Try with 'buffer=False'.")
    with paths["metadata"].open("r") as f:
        meta = json.load(f)
    device = str(agent.device) if getattr(agent, "device", None) is not None else "cpu"
    state: Dict[str, Any] = {}
    for key, t in load_file(paths["model_state"], device).items():
        module, param = key.split(".", 1)
        state.setdefault(module, {})[param] = t
    optim: Dict[str, Dict[int, Dict[str, torch.Tensor]]] = {}
    for key, t in load_file(paths["optim_state"], device).items():
        name, pid, k = key.split(".")
        optim.setdefault(name, {}).setdefault(int(pid), {})[k] = t
    for key, groups in meta.items():
        if key in ("model", "buffer"):
            continue
        state[key] = {"state": optim.get(key, {}), "param_groups": groups}
    for name in MODULE_NAMES:
        module = getattr(agent, name, None)
        if module is not None:
            module.load_state_dict(state)
    if buffer:
        agent.buffer = type(agent.buffer).load(paths["buffer_state"], meta["buffer"])
    return agent
