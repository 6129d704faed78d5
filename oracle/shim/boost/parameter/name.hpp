// stand-in (oracle/shim/README.md): the sliver of Boost.Parameter the reference uses --
//   BOOST_PARAMETER_NAME(x)            declares the keyword object _x (src/jpezy.hpp:346-362)
//   (_a = 1, _b = "s", ...)            builds an argument pack with operator,
//   pack[_a]                           reads a value back (src/jpezy.hpp:381-384)
#pragma once
#include <tuple>
#include <type_traits>
#include <utility>
namespace boost {
namespace parameter {
template <class Tag, class T>
struct tagged_arg {
    T value;
};
template <class Tag>
struct keyword {
    template <class T>
    constexpr tagged_arg<Tag, std::decay_t<T>> operator=(T&& v) const { return {std::forward<T>(v)}; }
};
template <class... Args>
struct arg_pack {
    std::tuple<Args...> args;
    template <class Tag>
    constexpr const auto& operator[](const keyword<Tag>&) const { return find<Tag, 0>(); }

private:
    template <class Tag, std::size_t I>
    constexpr const auto& find() const
    {
        static_assert(I < sizeof...(Args), "named argument not supplied");
        using A = std::tuple_element_t<I, std::tuple<Args...>>;
        if constexpr (is_tagged<Tag, A>::value) return std::get<I>(args).value;
        else return find<Tag, I + 1>();
    }
    template <class Tag, class A>
    struct is_tagged : std::false_type {};
    template <class Tag, class T>
    struct is_tagged<Tag, tagged_arg<Tag, T>> : std::true_type {};
};
template <class T1, class V1, class T2, class V2>
constexpr arg_pack<tagged_arg<T1, V1>, tagged_arg<T2, V2>> operator,(tagged_arg<T1, V1> a, tagged_arg<T2, V2> b)
{
    return {std::make_tuple(std::move(a), std::move(b))};
}
template <class... Args, class T2, class V2>
constexpr arg_pack<Args..., tagged_arg<T2, V2>> operator,(arg_pack<Args...> p, tagged_arg<T2, V2> b)
{
    return {std::tuple_cat(std::move(p.args), std::make_tuple(std::move(b)))};
}
}  // namespace parameter
}  // namespace boost
#define BOOST_PARAMETER_NAME(name)                                      \
    namespace tag {                                                     \
    struct name;                                                        \
    }                                                                   \
    static constexpr ::boost::parameter::keyword<tag::name> _##name{};
