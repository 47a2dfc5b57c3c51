"""Drop-in for the reference's perturbation pipeline module
(Training/functions/new_cvpr_train_behavior_things_pipeline.py, "NEW"): same public names and
signatures — the sweep drivers (uniform_sweep/clip_train_behavior_sweep.py,
length_experiments/clip_train_behavior_lengths.py) import ``run_behavioral_training`` from here and
run unmodified — with the arithmetic on libhba (sm_100a).  Implementation: functions/_pipeline_core.py.
"""
import csv
import os

import numpy as np
import torch
from torch.utils.data import DataLoader

from functions._pipeline_core import (  # noqa: F401  (re-exported reference surface)
    CLIPHBA, DoRALayer, NEW_HEADERS, Perturbation, SubsetWithIndices, ThingsDataset,
    ThingsInferenceDataset, append_csv_row, apply_dora_to_ViT, behavioral_RSA, build_model,
    count_trainable_parameters, describe_run, evaluate_model, load_clip_to_cpu,
    load_dataset_split_indices, load_random_states, make_optimizer, open_logger,
    replace_with_gaussian_noise, save_dora_parameters, save_random_states, seed_everything,
    select_device, setup_logger, shuffle_targets, switch_dora_layers, train_one_epoch,
    enable_trunk_cache, resident_loaders, CHECKPOINTS)
from functions.spose_dimensions import classnames66  # noqa: F401


def _prepare_results_csv(path, previous_path, resume_from_epoch, log, logger):
    """CSV bootstrap of NEW:797-834: append in place when resuming the same file, otherwise start a
    new file pre-populated with the rows (epoch <= resume_from_epoch) of the run resumed from."""
    same_file = previous_path == path and os.path.exists(path) and resume_from_epoch > 0
    if same_file:
        log("Resuming from existing CSV file - will append new epochs")
        try:
            with open(path, "r") as f:
                found = next(csv.reader(f), None)
            if found != NEW_HEADERS:
                log(f"Warning: CSV headers don't match. Expected {NEW_HEADERS}, found {found}")
        except Exception as e:  # noqa: BLE001
            if logger:
                logger.warning(f"Could not verify existing CSV file: {e}")
        return
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(NEW_HEADERS)
        if previous_path and resume_from_epoch > 0 and os.path.exists(previous_path):
            try:
                with open(previous_path, "r") as prev:
                    rows = csv.reader(prev)
                    next(rows, None)
                    for row in rows:
                        try:
                            if int(row[0]) <= resume_from_epoch:
                                w.writerow(row)
                        except Exception:  # noqa: BLE001
                            continue
            except Exception as e:  # noqa: BLE001
                if logger:
                    logger.warning(f"Could not pre-populate training CSV from {previous_path}: {e}")


def train_model(model, train_loader, test_loader, inference_loader, device, optimizer, criterion, epochs,
                training_res_path, training_run, perturb_length, perturb_seed, mean, std,
                perturb_distribution, perturb_type, logger=None, early_stopping_patience=5,
                checkpoint_path='clip_hba_model_cv.pth', dora_parameters_path='./dora_params',
                random_state_path='./random_states', dataloader_generator=None, resume_from_epoch=0,
                previous_training_res_path=None):
    """Epoch loop of NEW:782-1063 (perturbation window, per-epoch eval + RSA + CSV row + DoRA /
    random-state checkpoints, early stopping whose counter is frozen inside the window)."""
    model.train()
    log = logger.info if logger else print
    best_test_loss = 500000  # NEW:790
    epochs_no_improve = 0
    os.makedirs(dora_parameters_path, exist_ok=True)
    _prepare_results_csv(training_res_path, previous_training_res_path, resume_from_epoch, log, logger)
    perturb = Perturbation(perturb_type, training_run, perturb_length, perturb_seed,
                           perturb_distribution, mean, std)
    banner = {"random_target": "USING RANDOM TARGETS", "image_noise": "USING IMAGE NOISE",
              "label_shuffle": "USING SHUFFLED TARGETS", "uniform_images": "USING UNIFORM GRAYSCALE IMAGES"}
    done = {"used_random_targets": "RANDOM TARGETS WERE USED IN THIS EPOCH",
            "used_shuffled_targets": "SHUFFLED TARGETS WERE USED IN THIS EPOCH",
            "used_uniform_images": "UNIFORM GRAYSCALE IMAGES WERE USED IN THIS EPOCH",
            "used_image_noise": "GAUSSIAN NOISE WAS APPLIED TO IMAGES IN THIS EPOCH"}
    # (per-epoch checkpoints may be written by the background thread inside this scope; all of them are on
    # disk - or their error raised - when it is left)
    with CHECKPOINTS.deferred():
        for epoch in range(resume_from_epoch, epochs):
            flags = perturb.flags(epoch)
            if perturb.active(epoch):
                log("=" * 80)
                log(f"\n*** {banner[perturb_type]} FOR EPOCH {epoch+1} (Perturbation window: epochs "
                    f"{perturb.first+1}-{perturb.last+1}) ***")
                log("=" * 80)
                log(f"Perturbation seed: {perturb_seed}")
            avg_train_loss = train_one_epoch(model, train_loader, device, optimizer, criterion, epoch, epochs,
                                             perturb, log)
            avg_test_loss = evaluate_model(model, test_loader, device, criterion)
            log(f"Epoch {epoch+1}: Training Loss: {avg_train_loss:.4f}, Validation Loss: {avg_test_loss:.4f}")
            rho, p_value, _ = behavioral_RSA(model, inference_loader, device, logger=logger)
            log(f"Behavioral RSA Correlation & p-value: {rho:.4f}, {p_value:.4f}")
            model.train()
            for key, msg in done.items():
                if flags[key]:
                    log(f"*** {msg} ***")
            append_csv_row(training_res_path, [epoch + 1, avg_train_loss, avg_test_loss, rho, p_value,
                                               flags["used_random_targets"], flags["used_shuffled_targets"],
                                               flags["used_uniform_images"], flags["used_image_noise"]])
            save_dora_parameters(model, dora_parameters_path, epoch, logger=logger)
            log(f"DoRA parameters saved for epoch {epoch+1}")
            if dataloader_generator is not None:
                save_random_states(optimizer, epoch, random_state_path, dataloader_generator, logger=logger)
            if avg_test_loss < best_test_loss:
                best_test_loss, epochs_no_improve = avg_test_loss, 0
            elif not perturb.in_window(epoch):
                epochs_no_improve += 1
            if epochs_no_improve == early_stopping_patience:
                log("\n\n*********************************")
                log(f"Early stopping triggered at epoch {epoch+1}")
                log("*********************************\n\n")
                break


def run_behavioral_training(config):
    """NEW:1066-1227: one sweep condition (the sharding unit of the multi-GPU sweep runners)."""
    seed_everything(config['random_seed'])
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    logger = open_logger(config)
    dataset = ThingsDataset(csv_file=config['csv_file'], img_dir=config['img_dir'])
    embeddings = dataset.annotations.iloc[:, 1:].values.astype('float32')
    if config['perturb_distribution'] == 'normal':
        mean, std = 0, 1
    elif config['perturb_distribution'] == 'target':
        mean, std = np.mean(embeddings), np.std(embeddings)
    split_info = load_dataset_split_indices(config.get('baseline_split_indices_path'), logger=logger)
    train_dataset = SubsetWithIndices(dataset, split_info['train_indices'])
    test_dataset = SubsetWithIndices(dataset, split_info['test_indices'])
    logger.info("Using baseline dataset split")
    inference_dataset = ThingsInferenceDataset(inference_csv_file=config['inference_csv_file'],
                                               img_dir=config['img_dir'],
                                               RDM48_triplet_dir=config['RDM48_triplet_dir'])
    dataloader_generator = torch.Generator()
    dataloader_generator.manual_seed(config['random_seed'])
    device = select_device(config['cuda'])
    if config.get('hba_resident', True):
        # same batch order / generator consumption as the DataLoaders of NEW:1123-1126, images in HBM
        train_loader, test_loader, inference_loader = resident_loaders(
            config, dataset, train_dataset, test_dataset, inference_dataset, device, dataloader_generator,
            split_info['train_indices'], split_info['test_indices'])
    else:
        train_loader = DataLoader(train_dataset, batch_size=config['batch_size'], shuffle=True,
                                  generator=dataloader_generator)
        test_loader = DataLoader(test_dataset, batch_size=config['batch_size'], shuffle=False)
        inference_loader = DataLoader(inference_dataset, batch_size=config['batch_size'], shuffle=False)
    model = build_model(config, device, logger)
    training_run = config['training_run']
    resume_from_epoch = config.get('resume_from_epoch', 0)
    CHECKPOINTS.flush()   # checkpoints of an earlier condition in this process are complete before any is read
    if resume_from_epoch > 0 and config.get('resume_dora_parameters_path'):
        dora_path = os.path.join(config['resume_dora_parameters_path'],
                                 f"epoch{resume_from_epoch}_dora_params.pth")
    else:
        dora_path = os.path.join(config['baseline_dora_directory'], f"epoch{training_run - 1}_dora_params.pth")
    if dora_path and os.path.exists(dora_path) and training_run >= 1:
        model.load_state_dict(torch.load(dora_path), strict=False)
        logger.info(f"Loaded DoRA parameters from {dora_path}")
    else:
        logger.info("Using original DoRA parameters from model initialization")
    model.to(device)
    if config.get('hba_resident', True) and config.get('hba_trunk_cache', True):
        enable_trunk_cache(model, len(dataset) + len(inference_dataset))
    optimizer = make_optimizer(model, config['lr'])
    if resume_from_epoch > 0:
        prior = config.get('resume_random_state_path') or config.get('baseline_random_state_path')
        if prior:
            logger.info(f"Resuming from epoch {resume_from_epoch}")
            if load_random_states(prior, resume_from_epoch, optimizer=optimizer,
                                  dataloader_generator=dataloader_generator, logger=logger):
                logger.info(f"Successfully restored all random states from epoch {resume_from_epoch}")
            else:
                logger.warning("Could not load random states - starting with fresh random state")
        else:
            logger.warning("baseline_random_state_path not provided in config, cannot restore random states")
    describe_run(model, config, logger)
    train_model(model, train_loader, test_loader, inference_loader, device, optimizer,
                criterion=config['criterion'], epochs=config['epochs'],
                training_res_path=config['training_res_path'], training_run=training_run,
                perturb_length=config['perturb_length'], perturb_seed=config['perturb_seed'],
                mean=mean, std=std, perturb_distribution=config['perturb_distribution'],
                perturb_type=config['perturb_type'], logger=logger,
                early_stopping_patience=config['early_stopping_patience'],
                checkpoint_path=config['checkpoint_path'],
                dora_parameters_path=config['dora_parameters_path'],
                random_state_path=config.get('random_state_path', './random_states'),
                dataloader_generator=dataloader_generator, resume_from_epoch=resume_from_epoch,
                previous_training_res_path=config.get('previous_training_res_path'))
