"""Checkpoint I/O with the reference's on-disk layout, so weights/ directories written by either side load
in the other:

    trainer.py:143-203        model.safetensors            = model.state_dict()  (keys "unet.<module path>", 1239 tensors)
                              checkpoint-{step}/checkpoint.pt = {"model_state_dict", "optimizer_state_dict",
                                                                 "scheduler_state_dict", "rng_state"}
    trainer_peft.py:146-206   loras/checkpoint-{step}/     = peft `save_pretrained` (adapter_model.safetensors + adapter_config.json)
                                                             + checkpoint.pt without the model entry
                              merged_model.safetensors     = merge_and_unload(model).state_dict()

Function names and argument meaning follow the reference's own helpers.  Pure host code: no kernels involved.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Optional

import torch
from safetensors.torch import load_file, save_file

from . import lora


def save_model_sd(model: torch.nn.Module, project_dir: Path) -> None:
    """trainer.py:143-145."""
    project_dir = Path(project_dir)
    project_dir.mkdir(parents=True, exist_ok=True)
    save_file({k: v.detach().contiguous().cpu() for k, v in model.state_dict().items()}, str(project_dir / "model.safetensors"))


def load_model(model: torch.nn.Module, model_path: Path) -> None:
    """trainer_peft.py:146-152: a `checkpoint.pt` dict or a .safetensors file."""
    if str(model_path).endswith(".pt"):
        state_dict = torch.load(model_path, map_location="cpu", weights_only=False)["model_state_dict"]
    else:
        state_dict = load_file(str(model_path))
    model.load_state_dict(state_dict)


def save_checkpoint(model, optimizer, scheduler, current_step: int, project_dir: Path, is_nan: bool = False) -> Path:
    """trainer.py:148-178 (same directory name, file name and dictionary keys)."""
    checkpoint_dir = Path(project_dir) / f"checkpoint-{current_step + 1}{'-nan' if is_nan else ''}"
    checkpoint_dir.mkdir(parents=True, exist_ok=True)
    torch.save({
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict(),
        "scheduler_state_dict": scheduler.state_dict(),
        "rng_state": torch.get_rng_state(),
    }, checkpoint_dir / "checkpoint.pt")
    return checkpoint_dir


def load_checkpoint(model, optimizer, scheduler, checkpoint_path: Path, reset_steps: bool = False) -> int:
    """trainer.py:181-203, including the strict=False fallback when the model changed and the step parsed from the directory name."""
    checkpoint_path = Path(checkpoint_path)
    device = next(model.parameters()).device
    checkpoint = torch.load(checkpoint_path / "checkpoint.pt", map_location=device, weights_only=False)
    try:
        model.load_state_dict(checkpoint["model_state_dict"])
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    except (RuntimeError, ValueError):     # RuntimeError: model keys changed (trainer.py:191); ValueError: optimizer group sizes differ
        model.load_state_dict(checkpoint["model_state_dict"], strict=False)
    if not reset_steps:
        scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    torch.set_rng_state(checkpoint["rng_state"].cpu())
    return 0 if reset_steps else int(checkpoint_path.stem.split("-")[1])


def get_latest_checkpoint(project_dir: Path, sub: str = "") -> Optional[Path]:
    """trainer.py / trainer_peft.py:155-158: highest-numbered `checkpoint-*` directory (under `loras/` for adapters)."""
    root = Path(project_dir) / sub if sub else Path(project_dir)
    ckpts = [p for p in root.rglob("checkpoint-*") if p.is_dir() and p.stem.split("-")[1].isdigit()]
    ckpts.sort(key=lambda p: int(p.stem.split("-")[1]))
    return ckpts[-1] if ckpts else None


# ------------------------------------------------------------------------------------------------ adapters (peft layout)
def save_adapter(model: torch.nn.Module, checkpoint_dir: Path, r: int = 32, lora_alpha: int = 32, use_dora: bool = True,
                 target_modules=lora.DEFAULT_TARGETS) -> None:
    """What `PeftModel.save_pretrained(dir)` writes (peft 0.12.0): adapter_model.safetensors with
    `base_model.model.<path>.lora_{A,B}.weight` / `.lora_magnitude_vector.weight` keys + adapter_config.json."""
    checkpoint_dir = Path(checkpoint_dir)
    checkpoint_dir.mkdir(parents=True, exist_ok=True)
    sd = {k: v.contiguous().cpu() for k, v in lora.adapter_state_dict(model).items()}
    save_file(sd, str(checkpoint_dir / "adapter_model.safetensors"))
    cfg = {"peft_type": "LORA", "r": r, "lora_alpha": lora_alpha, "use_dora": use_dora, "target_modules": list(target_modules),
           "lora_dropout": 0.0, "bias": "none", "fan_in_fan_out": False, "init_lora_weights": True, "task_type": None,
           "base_model_name_or_path": None, "inference_mode": True}
    (checkpoint_dir / "adapter_config.json").write_text(json.dumps(cfg, indent=2))


def load_adapter(model: torch.nn.Module, checkpoint_dir: Path) -> dict:
    """Load an adapter directory written by `save_adapter` or by peft into a model prepared with `lora.inject_adapters`."""
    checkpoint_dir = Path(checkpoint_dir)
    lora.load_adapter_state_dict(model, load_file(str(checkpoint_dir / "adapter_model.safetensors")))
    cfg_path = checkpoint_dir / "adapter_config.json"
    return json.loads(cfg_path.read_text()) if cfg_path.exists() else {}


def save_peft_checkpoint(model, optimizer, scheduler, current_step: int, project_dir: Path, **adapter_cfg) -> Path:
    """trainer_peft.py:167-190."""
    checkpoint_dir = Path(project_dir) / "loras" / f"checkpoint-{current_step + 1}"
    save_adapter(model, checkpoint_dir, **adapter_cfg)
    torch.save({
        "optimizer_state_dict": optimizer.state_dict(),
        "scheduler_state_dict": scheduler.state_dict(),
        "rng_state": torch.get_rng_state(),
    }, checkpoint_dir / "checkpoint.pt")
    return checkpoint_dir


def load_peft_checkpoint(optimizer, scheduler, checkpoint_path: Path, reset_steps: bool) -> int:
    """trainer_peft.py:193-206."""
    checkpoint_path = Path(checkpoint_path)
    checkpoint = torch.load(checkpoint_path / "checkpoint.pt", weights_only=False)
    optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    if not reset_steps:
        scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    torch.set_rng_state(checkpoint["rng_state"])
    return 0 if reset_steps else int(checkpoint_path.stem.split("-")[1])


def save_merged_model_sd(model: torch.nn.Module, project_dir: Path) -> None:
    """trainer_peft.py:161-164: merge_and_unload, then the plain 1239-key state_dict."""
    merged = lora.merge_and_unload(model)
    project_dir = Path(project_dir)
    project_dir.mkdir(parents=True, exist_ok=True)
    save_file({k: v.detach().contiguous().cpu() for k, v in merged.state_dict().items()}, str(project_dir / "merged_model.safetensors"))
