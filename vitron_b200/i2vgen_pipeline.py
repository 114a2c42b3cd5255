"""i2vgen-xl image-to-video sampling on the vitron_b200 modules: the inner part of `inference_i2vgen_entrance.worker`
(modules/i2vgen-xl/tools/inferences/inference_i2vgen_entrance.py:118-209) as one callable.

    clip_encoder(text="") / (text=negative_prompt)                      :120-123   -> zero_y, zero_y_negative
    y_visual, y_text, y_words = clip_encoder(image=vit image, text=caption)  :166-168
    local_image = autoencoder.encode_firsr_stage(image, scale_factor) repeated over max_frames   :171-174
    noise [1, 4, F, H/8, W/8]; model_kwargs = [cond, uncond]; diffusion.ddim_sample_loop(...)    :186-198
    video = autoencoder.decode(latents / scale_factor) in chunks of decoder_bs frames             :200-209

Decoding the input image, its PIL BOX-resize transforms (`CenterCropWide`, utils/transforms.py:163-183), tokenisation
and mp4 writing stay on the host and are not part of this module: inputs are the transformed tensors / token ids.
Every tensor between the stages stays on the device; the DDIM loop replays one CUDA graph per step.
"""
import torch

from .unet_i2vgen import DiffusionDDIM, GraphedCFGDenoiser

BF16 = torch.bfloat16


class I2VGenXLPipeline:
    def __init__(self, unet, autoencoder, clip_encoder, diffusion=None, scale_factor=0.18215, max_frames=16, guide_scale=9.0,
                 ddim_timesteps=50, decoder_bs=8, use_zero_infer=True, target_fps=16, use_graph=True):
        self.unet, self.autoencoder, self.clip_encoder = unet, autoencoder, clip_encoder
        self.diffusion = diffusion or DiffusionDDIM(
            schedule="cosine", schedule_param=dict(num_timesteps=1000, cosine_s=0.008, zero_terminal_snr=True), mean_type="v",
            var_type="fixed_small")                                              # tools/modules/config.py:55-68
        self.scale_factor, self.max_frames, self.guide_scale = scale_factor, max_frames, guide_scale
        self.ddim_timesteps, self.decoder_bs, self.use_zero_infer, self.target_fps = ddim_timesteps, decoder_bs, use_zero_infer, target_fps
        self.use_graph = use_graph
        self.device = unet.device if hasattr(unet, "device") else torch.device("cuda")
        self._den = None      # the captured CFG evaluation, reused across videos of the same shape (rebind)

    @torch.no_grad()
    def __call__(self, image_vit, image_vae, tokens, negative_tokens, noise=None, generator=None, posterior_noise=None):
        """image_vit [1, 3, 224, 224] (vit_trans output), image_vae [1, 3, H, W] (train_trans output), tokens /
        negative_tokens [1, 77] open_clip ids -> video [1, 3, F, H, W] fp32 (normalised like train_trans)."""
        dev = self.device
        y_visual, _, y_words = self.clip_encoder(image=image_vit.to(dev), text=tokens.to(dev))
        y_visual = y_visual.unsqueeze(1)                                           # [1, 1, 1024]
        _, _, zero_y_negative = self.clip_encoder(text=negative_tokens.to(dev))
        local = self.autoencoder.encode_firsr_stage(image_vae.to(dev), self.scale_factor, noise=posterior_noise)
        local_image = local.unsqueeze(2).repeat_interleave(repeats=self.max_frames, dim=2)       # [1, 4, F, h, w]
        b, _, h, w = local.shape
        if noise is None:
            noise = torch.randn((b, 4, self.max_frames, h, w), generator=generator).to(dev)
        fps = torch.tensor([self.target_fps], dtype=torch.long, device=dev)
        infer_img = torch.zeros_like(y_visual) if self.use_zero_infer else None      # black_image_feature :124
        model_kwargs = [dict(y=y_words, image=y_visual, local_image=local_image, fps=fps),
                        dict(y=zero_y_negative, image=infer_img, local_image=local_image, fps=fps)]
        noise = noise.to(dev)
        model = self.unet
        if self.use_graph:
            if self._den is not None and tuple(self._den.xt.shape) == tuple(noise.shape):
                self._den.rebind(*model_kwargs)                       # same graph, new conditioning
            else:
                own = {}                                               # the denoiser owns static copies (local_image shared)

                def static(d):
                    out = {}
                    for k, v in d.items():
                        if torch.is_tensor(v):
                            own.setdefault(id(v), v.clone())
                            out[k] = own[id(v)]
                        else:
                            out[k] = v
                    return out
                self._den = GraphedCFGDenoiser(self.unet, static(model_kwargs[0]), static(model_kwargs[1]), self.guide_scale, noise,
                                               torch.zeros((b,), dtype=torch.long, device=dev))
            model = self._den
        latents = self.diffusion.ddim_sample_loop(noise=noise, model=model, model_kwargs=model_kwargs,
                                                  guide_scale=self.guide_scale, ddim_timesteps=self.ddim_timesteps, eta=0.0)
        latents = (1.0 / self.scale_factor) * latents                               # :200
        frames = latents.permute(0, 2, 1, 3, 4).reshape(b * self.max_frames, 4, h, w)   # 'b c f h w -> (b f) c h w'
        chunk = min(self.decoder_bs, frames.shape[0])
        dec = [self.autoencoder.decode(fr) for fr in torch.chunk(frames, frames.shape[0] // chunk, dim=0)]
        video = torch.cat(dec, dim=0)
        return video.reshape(b, self.max_frames, *video.shape[1:]).permute(0, 2, 1, 3, 4).contiguous()
