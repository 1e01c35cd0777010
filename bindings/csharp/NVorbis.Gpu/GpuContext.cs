using System;
using System.Runtime.InteropServices;

namespace NVorbis.Gpu
{
    /// <summary>One vpz_ctx = one GPU. All calls on a context from one thread at a time (the reference's decoders are
    /// not thread-safe either: Mapping.cs:17, Residue0.cs:22); one host thread + context per GPU shards streams.</summary>
    public sealed class GpuContext : IDisposable
    {
        static readonly Lazy<GpuContext> s_default = new(() => new GpuContext(-1));
        public static GpuContext Default => s_default.Value;
        public static int DeviceCount => Vpz.vpz_device_count();

        internal IntPtr Handle { get; private set; }

        public GpuContext(int device = -1)
        {
            int rc = Vpz.vpz_ctx_create(device, out IntPtr h);
            Vpz.Check(rc, IntPtr.Zero);     // VPZ_E_NO_DEVICE -> PlatformNotSupportedException: there is no CPU path
            Handle = h;
        }

        /// <summary>Tunables of include/vpz.h ("l1_bits", "ola_chunk", "bulk_group", "host_threads", "gpu_scan", ...).</summary>
        public void Set(string key, int value) => Vpz.Check(Vpz.vpz_ctx_set(Handle, key, value), Handle);

        public long KernelLaunches => Vpz.vpz_ctx_kernel_launches(Handle);
        public static string Version => Marshal.PtrToStringUTF8(Vpz.vpz_version()) ?? "";

        /// <summary>Pinned host memory (full PCIe rate for the bulk calls' destination).</summary>
        public unsafe PinnedBuffer<T> AllocPinned<T>(long count) where T : unmanaged
        {
            IntPtr p = Vpz.vpz_host_alloc((nuint)(count * sizeof(T)));
            if (p == IntPtr.Zero) throw new OutOfMemoryException();
            return new PinnedBuffer<T>(p, count);
        }

        public void Dispose()
        {
            if (Handle != IntPtr.Zero) Vpz.vpz_ctx_destroy(Handle);
            Handle = IntPtr.Zero;
        }
    }

    public sealed unsafe class PinnedBuffer<T> : IDisposable where T : unmanaged
    {
        public IntPtr Pointer { get; private set; }
        public long Length { get; }
        internal PinnedBuffer(IntPtr p, long n) { Pointer = p; Length = n; }
        public Span<T> Slice(long start, int count) => new((T*)Pointer + start, count);
        public void Dispose()
        {
            if (Pointer != IntPtr.Zero) Vpz.vpz_host_free(Pointer);
            Pointer = IntPtr.Zero;
        }
    }
}
