"""ORACLE - test infrastructure only.

Nothing under vit-project_b200/ imports this package; only tests/, `__graft_entry__.smoke()` and the CPU legs of
bench.py do, and only as the checker.  The product has no CPU path: `import hba` needs the compiled libhba.so and
every front end refuses non-CUDA tensors.

  restatements (CPU, each citing the reference lines it follows)
    clip_ref.py          the un-vendored CLIP towers (OpenAI ViT variants; "parity unpinned", cross-checked against
                         transformers.CLIPModel)
    dora_ref.py          DoRALayer / apply_dora_to_ViT / switch_dora_layers / CLIPHBA       NEW:268-304, 407-544
    ops_ref.py           formulas of the individual kernels (attention, LayerNorm, cosine head, RSA tail)
    vit_ref.py           timm ViT-B/16 (cross-checked against torchvision.models.vit_b_16)   VIT:283
    vit_measure_ref.py   the ViT epoch / measurement functions                               VIT:89-244, MEAS:36-555
    analysis_ref.py      the figure notebooks' tables, row by row
    libhba_ref.py        the C-ABI of include/hba.h itself, entry point by entry point, on host memory - lets the
                         tests run hba.engine / hba.vit / the pipelines on CPU tensors (`emulated_device()`)
  the reference EXECUTED here (where /root/reference is mounted; the outputs are committed under tests/golden/)
    ref_loader.py        imports the reference's own pipeline modules with the plug-in `clip` stubbed
    make_golden.py, make_analysis_golden.py, make_vit_measure_golden.py   golden generators
    clip_train_exec.py   `train_model` of NEW and BASE vs this repo's epoch loops (stand-in network, bit for bit)
    clip_pipeline_exec.py  `run_behavioral_training` of BASE and NEW vs this repo's, from the files to the CSV
    vit_measure_exec.py  the ViT scripts' epoch / measurement functions at world sizes 1 and 2
    run_notebook.py      the figure notebooks' cells on the reference's Data/ tree
  synth.py               seeded synthetic inputs shared by the generators and the tests
"""
