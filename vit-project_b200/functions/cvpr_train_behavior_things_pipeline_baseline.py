"""Drop-in for the reference's baseline pipeline module
(Training/functions/cvpr_train_behavior_things_pipeline_baseline.py, "BASE"): the un-perturbed run
that creates the 80/20 split and the per-epoch checkpoints every sweep condition resumes from.
Same public names and signatures; arithmetic on libhba.  Implementation: functions/_pipeline_core.py.
"""
import os

import torch
from torch.utils.data import DataLoader, random_split

from functions._pipeline_core import (  # noqa: F401  (re-exported reference surface)
    BASE_HEADERS, CLIPHBA, DoRALayer, ThingsDataset, ThingsInferenceDataset, append_csv_row,
    apply_dora_to_ViT, build_model, count_trainable_parameters, describe_run, evaluate_model,
    load_clip_to_cpu, make_optimizer, open_logger, save_random_states, seed_everything, select_device,
    setup_logger, switch_dora_layers, train_one_epoch, enable_trunk_cache, resident_loaders, CHECKPOINTS)
from functions._pipeline_core import behavioral_RSA as _behavioral_RSA
from functions.spose_dimensions import classnames66  # noqa: F401
from src.models.clip_hba_utils import save_dora_parameters


def behavioral_RSA(model, inference_loader, device):
    """BASE:490-575 (no logger argument in this module)."""
    return _behavioral_RSA(model, inference_loader, device)


def train_model(model, train_loader, test_loader, inference_loader, device, optimizer, criterion, epochs,
                training_res_path, logger=None, early_stopping_patience=5,
                checkpoint_path='clip_hba_model_cv.pth', dora_parameters_path='./dora_params',
                random_state_path='./random_states', dataloader_generator=None, vision_layers=1,
                transformer_layers=1):
    """BASE:612-704: initial evaluation, then the plain epoch loop."""
    model.train()
    log = logger.info if logger else print
    log("*********************************")
    log("Evaluating initial model")
    best_test_loss = evaluate_model(model, test_loader, device, criterion)
    rho0, p0, _ = behavioral_RSA(model, inference_loader, device)
    log(f"Initial Validation Loss: {best_test_loss:.4f}")
    log(f"Initial Behavioral RSA Correlation & p-value: {rho0:.4f}, {p0:.4f}")
    log("*********************************\n")
    model.train()
    os.makedirs(dora_parameters_path, exist_ok=True)
    os.makedirs(os.path.dirname(training_res_path) or ".", exist_ok=True)
    with open(training_res_path, 'w', newline='') as f:
        f.write(",".join(BASE_HEADERS) + "\r\n")
    epochs_no_improve = 0
    # (per-epoch checkpoints may be written by the background thread inside this scope; all of them are on
    # disk - or their error raised - when it is left)
    with CHECKPOINTS.deferred():
        for epoch in range(epochs):
            avg_train_loss = train_one_epoch(model, train_loader, device, optimizer, criterion, epoch, epochs,
                                             None, log)
            avg_test_loss = evaluate_model(model, test_loader, device, criterion)
            log(f"Epoch {epoch+1}: Training Loss: {avg_train_loss:.4f}, Validation Loss: {avg_test_loss:.4f}")
            rho, p_value, _ = behavioral_RSA(model, inference_loader, device)
            log(f"Behavioral RSA Correlation & p-value: {rho:.4f}, {p_value:.4f}")
            model.train()
            append_csv_row(training_res_path, [epoch + 1, avg_train_loss, avg_test_loss, rho, p_value])
            save_random_states(optimizer, epoch, random_state_path, dataloader_generator, logger=logger)
            save_dora_parameters(model, dora_parameters_path, epoch, vision_layers, transformer_layers,
                                 log_fn=log)
            log(f"DoRA parameters saved for epoch {epoch+1}")
            if avg_test_loss < best_test_loss:
                best_test_loss, epochs_no_improve = avg_test_loss, 0
            else:
                epochs_no_improve += 1
            if epochs_no_improve == early_stopping_patience:
                log("\n\n*********************************")
                log(f"Early stopping triggered at epoch {epoch+1}")
                log("*********************************\n\n")
                break


def run_behavioral_training(config):
    """BASE:707-823: seed -> 80/20 random_split (saved as dataset_split_indices.pth) -> model + DoRA
    -> AdamW -> train_model."""
    seed_everything(config['random_seed'])
    logger = open_logger(config)
    dataset = ThingsDataset(csv_file=config['csv_file'], img_dir=config['img_dir'])
    train_size = int(config['train_portion'] * len(dataset))
    train_dataset, test_dataset = random_split(dataset, [train_size, len(dataset) - train_size])
    os.makedirs(config['random_state_path'], exist_ok=True)
    split_file = os.path.join(config['random_state_path'], 'dataset_split_indices.pth')
    torch.save({'train_indices': list(train_dataset.indices), 'test_indices': list(test_dataset.indices),
                'random_seed': config['random_seed'], 'train_portion': config['train_portion']}, split_file)
    logger.info(f"Dataset split indices saved: {split_file}")
    inference_dataset = ThingsInferenceDataset(inference_csv_file=config['inference_csv_file'],
                                               img_dir=config['img_dir'],
                                               RDM48_triplet_dir=config['RDM48_triplet_dir'])
    dataloader_generator = torch.Generator()
    dataloader_generator.manual_seed(config['random_seed'])
    device = select_device(config['cuda'])
    if config.get('hba_resident', True):
        train_loader, test_loader, inference_loader = resident_loaders(
            config, dataset, train_dataset, test_dataset, inference_dataset, device, dataloader_generator,
            list(train_dataset.indices), list(test_dataset.indices))
    else:
        train_loader = DataLoader(train_dataset, batch_size=config['batch_size'], shuffle=True,
                                  generator=dataloader_generator)
        test_loader = DataLoader(test_dataset, batch_size=config['batch_size'], shuffle=False)
        inference_loader = DataLoader(inference_dataset, batch_size=config['batch_size'], shuffle=False)
    model = build_model(config, device, logger)
    model.to(device)
    if config.get('hba_resident', True) and config.get('hba_trunk_cache', True):
        enable_trunk_cache(model, len(dataset) + len(inference_dataset))
    optimizer = make_optimizer(model, config['lr'])
    describe_run(model, config, logger)
    train_model(model, train_loader, test_loader, inference_loader, device, optimizer, config['criterion'],
                config['epochs'], config['training_res_path'], logger=logger,
                early_stopping_patience=config['early_stopping_patience'],
                checkpoint_path=config['checkpoint_path'],
                dora_parameters_path=config['dora_parameters_path'],
                random_state_path=config['random_state_path'], dataloader_generator=dataloader_generator,
                vision_layers=config['vision_layers'], transformer_layers=config['transformer_layers'])
